// The reference's own unit tests (src/lib.rs:420-738) re-expressed on the C++ host mirror (host/bjj.hpp).
// Built by tests/test_cpp_host.py; runs on the GPU box (`-m gpu`).  Exit code 0 = all passed.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../host/bjj.hpp"

using namespace bjj_host;

static U256 from_hex_be(const std::string& h) {      // big-endian hex (as printed by Fr's Display) -> LE bytes
    U256 r{};
    std::string s = h;
    while (s.size() < 64) s = "0" + s;
    for (int i = 0; i < 32; i++) r[31 - i] = (uint8_t)std::stoi(s.substr(2 * i, 2), nullptr, 16);
    return r;
}
static std::array<uint8_t, 32> from_hex_le(const std::string& h) {
    std::array<uint8_t, 32> r{};
    for (int i = 0; i < 32; i++) r[i] = (uint8_t)std::stoi(h.substr(2 * i, 2), nullptr, 16);
    return r;
}
#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main() {
    // P of the reference tests, x/y as 64-digit hex
    Point p{from_hex_be("274dbce8d15179969bc0d49fa725bddf9de555e0ba6a693c6adb52fc9ee7a82c"),
            from_hex_be("05ce98c61b05f47fe2eae9a542bd99f6b2e78246231640b54595febfd51eb853")};
    // test_add_same_point (src/lib.rs:420-459)
    Point d = p.projective().add(p.projective()).affine();
    CHECK(d.x == from_hex_be("0f3c160e26fc96c347dd9e705eb5a3e8d661502728609ff95b3b889296901ab5"));
    CHECK(d.y == from_hex_be("09979273078b5c735585107619130e62e315c5cafe683a064f79dfed17eb14e1"));
    Point d2 = p.mul_scalar(u256_from_u64(2));
    CHECK(d.equals(d2));
    // test_mul_scalar (src/lib.rs:501-552): 3P == P + P + P
    Point m3 = p.mul_scalar(u256_from_u64(3));
    Point a3 = p.projective().add(p.projective()).add(p.projective()).affine();
    CHECK(m3.equals(a3));
    // test_point_compress_decompress (src/lib.rs:574-594)
    auto c = p.compress();
    CHECK(c == from_hex_le("53b81ed5bffe9545b54016234682e7b2f699bd42a5e9eae27ff4051bc698ce85"));
    CHECK(decompress_point(c).equals(p));
    bool threw = false;
    try {
        decompress_point(from_hex_le("0100000000000000000000000000000000000000000000000000000000000000"));
    } catch (const std::invalid_argument& e) {
        threw = std::string(e.what()) == "not a mod p square";
    }
    CHECK(threw);
    // test_circomlib_testvector (src/lib.rs:688-738)
    PrivateKey sk = PrivateKey::import(std::vector<uint8_t>{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1});
    Point pk = sk.public_key();
    CHECK(pk.x == from_hex_be("1d5ac1f31407018b7d413a4f52c8f74463b30e6ac2238220ad8b254de4eaa3a2"));
    CHECK(pk.y == from_hex_be("1e1de8a908826c3f9ac2e0ceee929ecd0caf3b99b3ef24523aaab796a6f733c4"));
    U256 msg{};
    for (int i = 0; i < 10; i++) msg[i] = (uint8_t)i;
    Signature sig = sk.sign(msg);
    CHECK(sig.r_b8.x == from_hex_be("192b4e51adf302c8139d356d0e08e2404b5ace440ef41fc78f5c4f2428df0765"));
    CHECK(sig.r_b8.y == from_hex_be("2202bebcf57b820863e0acc88970b6ca7d987a0d513c2ddeb42e3f5d31b4eddf"));
    CHECK(verify(pk, sig, msg));
    U256 msg2 = msg;
    msg2[0] ^= 1;
    CHECK(!verify(pk, sig, msg2));
    // test_signature_compress_decompress (src/lib.rs:656-675)
    Signature sig2 = decompress_signature(sig.compress());
    CHECK(sig2.r_b8.equals(sig.r_b8) && sig2.s == sig.s);
    // batch entry points
    auto oks = verify_batch({pk, pk}, {sig, sig2}, {msg, msg2});
    CHECK(oks[0] == 1 && oks[1] == 0);
    auto pts = mul_scalar_batch({p, p}, {u256_from_u64(3), u256_from_u64(0)});
    CHECK(pts[0].equals(m3) && pts[1].x == u256_from_u64(0) && pts[1].y == u256_from_u64(1));
    CHECK(public_batch({sk})[0].equals(pk));
    auto dec = decompress_batch({c, from_hex_le("0100000000000000000000000000000000000000000000000000000000000000")});
    CHECK(dec[0].status == 0 && dec[0].point.equals(p) && dec[1].status == BJJ_STATUS_NOT_SQUARE);
    // multi-GPU sharding (works with one device too)
    MultiGpu mg;
    std::vector<uint8_t> keys(32 * 100), rx(32 * 100), ry(32 * 100);
    for (size_t i = 0; i < keys.size(); i++) keys[i] = (uint8_t)(i * 37 + 11);
    std::memcpy(keys.data(), sk.key.data(), 32);
    mg.public_batch(100, keys.data(), rx.data(), ry.data());
    CHECK(std::memcmp(rx.data(), pk.x.data(), 32) == 0);
    std::printf("host mirror ok (%zu device(s))\n", mg.devices());
    return 0;
}
