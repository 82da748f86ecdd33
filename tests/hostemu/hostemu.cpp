// TEST HARNESS ONLY -- never shipped, never loaded by the product.
//
// Compiles the device headers (babyjubjub-rs_b200/csrc/*.cuh) for the HOST with BJJ_HOST_EMU: the PTX
// carry chains are replaced by their 64-bit C emulation, everything else (limb schedules, curve
// formulas, recodings, table logic, Poseidon, square root) is the same source the GPU kernels
// compile.  `pytest -m "not gpu"` uses it to validate that logic against the oracle in a container
// without a GPU.  It is deliberately slow (one lane at a time, no threads).
#define BJJ_HOST_EMU 1
#include <stdlib.h>
#include <vector>
#include "../../babyjubjub-rs_b200/csrc/lanes.cuh"

using namespace bjj;

static CombEntry* g_comb = nullptr;
static uint8_t* g_comb_valid = nullptr;
static bool g_split = true;

// comb entries are built on demand (the device builds all 524,290 in k_comb_build; here a test touches a few)
namespace bjj {
void bjj_emu_need_comb_entry(const CombEntry* comb, int w, int j) {
    const size_t idx = (size_t)w * BJJ_COMB_ENTRIES + j;
    if (comb != g_comb || g_comb_valid[idx]) return;
    g_comb_valid[idx] = 1;
    comb_build_entry(g_comb, w, j);
}
}  // namespace bjj
static std::vector<U128> g_table(2 * BJJ_TABLE_U128_PER_LANE);

static std::vector<uint32_t> g_list, g_list2;
static uint32_t g_count, g_count2;
static ExactQueue exact_queue2(size_t n) {
    g_list2.assign(n + 1, 0);
    g_count2 = 0;
    ExactQueue q;
    q.count = &g_count2;
    q.list = g_list2.data();
    return q;
}
static ExactQueue exact_queue(size_t n) {
    g_list.assign(n + 1, 0);
    g_count = 0;
    ExactQueue q;
    q.count = &g_count;
    q.list = g_list.data();
    return q;
}

static std::vector<uint8_t> g_scr;
static ProjScratch proj_scratch(size_t n) {
    g_scr.assign(4 * 32 * (n + 1), 0);
    ProjScratch s;
    s.x = g_scr.data();
    s.y = s.x + 32 * n;
    s.z = s.y + 32 * n;
    s.p = s.z + 32 * n;
    return s;
}
// the batched affine pass with T "threads" (T = 3 exercises several strides and ragged tails)
static void batch_affine(const ProjScratch& s, uint8_t* rx, uint8_t* ry, size_t n) {
    const size_t T = 3;
    for (size_t t = 0; t < T; t++) batch_affine_strided(s, rx, ry, n, t, T);
}

static ScalarPark host_park() {
    static uint32_t words[BJJ_PARK_WORDS];
    return ScalarPark{words, 1};
}
static LaneTable lane_table(int which = 0) {
    LaneTable t;
    t.base = g_table.data() + (size_t)which * BJJ_TABLE_U128_PER_LANE;
    t.stride = 1;
    t.slot = 0;
    return t;
}

extern "C" {

void emu_init() {
    if (g_comb) return;
    g_comb = (CombEntry*)calloc(BJJ_COMB_TOTAL, sizeof(CombEntry));     // pages are committed on first touch
    g_comb_valid = (uint8_t*)calloc(BJJ_COMB_TOTAL, 1);
}

uint32_t emu_fr_op(int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) lane_fr_op(op, a, b, out, i, flags);
    return flags;
}

uint32_t emu_fr_dot6(size_t n, const uint8_t* a /*6 arrays concatenated*/, const uint8_t* b, uint8_t* out) {
    // out_i = sum_p a_p[i] * b_p[i]   (exercises fr_dot<6>)
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) {
        Fr A[6], B[6], r;
        for (int p = 0; p < 6; p++) {
            load_fr(A[p], a + 32 * n * p, i, flags);
            fr_reduce(A[p]);
            load_fr(B[p], b + 32 * n * p, i, flags);
        }
        fr_dot<6>(r, A, B);
        store_fr(out, i, r);
    }
    return flags;
}

uint32_t emu_add(size_t n, const uint8_t* px, const uint8_t* py, const uint8_t* pz, const uint8_t* qx,
                 const uint8_t* qy, const uint8_t* qz, uint8_t* rx, uint8_t* ry, uint8_t* rz) {
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) lane_add(px, py, pz, qx, qy, qz, rx, ry, rz, i, flags);
    return flags;
}

uint32_t emu_affine(size_t n, const uint8_t* px, const uint8_t* py, const uint8_t* pz, uint8_t* rx, uint8_t* ry) {
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) lane_affine(px, py, pz, rx, ry, i, flags);
    return flags;
}

uint32_t emu_mul_scalar(size_t n, const uint8_t* px, const uint8_t* py, const uint8_t* k, uint8_t* rx, uint8_t* ry) {
    uint32_t flags = 0;
    ExactQueue q = exact_queue(n);
    ProjScratch scr = proj_scratch(n);
    for (size_t i = 0; i < n; i++) lane_mul_scalar(px, py, k, scr, i, lane_table(), q, flags);
    batch_affine(scr, rx, ry, n);
    for (uint32_t j = 0; j < g_count; j++) lane_mul_scalar_exact(px, py, k, 8, rx, ry, g_list[j]);
    return flags;
}

void emu_fixed_base(size_t n, const uint8_t* k, uint8_t* rx, uint8_t* ry) {
    emu_init();
    ProjScratch scr = proj_scratch(n);
    for (size_t i = 0; i < n; i++) lane_fixed_base(k, scr, i, g_comb);
    batch_affine(scr, rx, ry, n);
}

void emu_public(size_t n, const uint8_t* key, uint8_t* rx, uint8_t* ry) {
    emu_init();
    ProjScratch scr = proj_scratch(n);
    for (size_t i = 0; i < n; i++) lane_public(key, scr, i, g_comb);
    batch_affine(scr, rx, ry, n);
}

void emu_scalar_key(size_t n, const uint8_t* key, uint8_t* out) {
    for (size_t i = 0; i < n; i++) lane_scalar_key(key, out, i);
}

void emu_sign(size_t n, const uint8_t* key, const uint8_t* msg, uint8_t* rx, uint8_t* ry, uint8_t* s32, uint8_t* status) {
    emu_init();
    for (size_t i = 0; i < n; i++) lane_sign(key, msg, rx, ry, s32, status, i, g_comb);
}

uint32_t emu_compress(size_t n, const uint8_t* px, const uint8_t* py, uint8_t* out) {
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) lane_compress(px, py, out, i, flags);
    return flags;
}

// batched decompression exactly as the kernels sequence it: prepare -> batched inverse -> finish
static void batch_inverse(const ProjScratch& s, size_t n) {
    const size_t T = 3;
    for (size_t t = 0; t < T; t++) batch_inverse_strided(s, n, t, T);
}

void emu_decompress(size_t n, const uint8_t* in, uint8_t* rx, uint8_t* ry, uint8_t* status) {
    ProjScratch scr = proj_scratch(n);
    for (size_t i = 0; i < n; i++) lane_decompress_prepare(in, 1, 0, scr, i, i);
    batch_inverse(scr, n);
    for (size_t i = 0; i < n; i++) lane_decompress_finish(in, 1, 0, scr, i, rx, ry, status, i, false);
}

uint32_t emu_poseidon(int n_inputs, size_t n, const uint8_t* const* in, uint8_t* out) {
    uint32_t flags = 0;
    for (size_t i = 0; i < n; i++) {
        switch (n_inputs) {
            case 1: lane_poseidon<2>(in, out, i, flags); break;
            case 2: lane_poseidon<3>(in, out, i, flags); break;
            case 3: lane_poseidon<4>(in, out, i, flags); break;
            case 4: lane_poseidon<5>(in, out, i, flags); break;
            case 5: lane_poseidon<6>(in, out, i, flags); break;
            case 6: lane_poseidon<7>(in, out, i, flags); break;
            default: return 0x80000000u;
        }
    }
    return flags;
}

uint32_t emu_verify(size_t n, const uint8_t* r8x, const uint8_t* r8y, const uint8_t* s, const uint8_t* ax,
                    const uint8_t* ay, const uint8_t* msg, uint8_t* ok) {
    emu_init();
    uint32_t flags = 0;
    ExactQueue qa = exact_queue(n), qr = exact_queue2(n);
    std::vector<uint8_t> hm(4 * 32 * (n + 1));
    for (size_t i = 0; i < n; i++)
        lane_verify_hash(r8x, r8y, ax, ay, msg, s, 1, 0, nullptr, hm.data(), n, ok, i, true, qa, qr, flags, BJJ_MODE_EDDSA, g_split, nullptr);
    if (g_split)
        for (size_t i = 0; i < n; i++) lane_verify_split(s, 1, 0, hm.data(), n, ok, i);
    for (size_t i = 0; i < n; i++) lane_verify_ec(r8x, r8y, ax, ay, hm.data(), n, ok, i, lane_table(0), lane_table(1), g_comb, BJJ_MODE_EDDSA, host_park());
    for (uint32_t j = 0; j < g_count; j++) lane_verify_exact<true>(r8x, r8y, s, ax, ay, hm.data(), ok, g_list[j], g_comb, BJJ_MODE_EDDSA);
    for (uint32_t j = 0; j < g_count2; j++) lane_verify_exact<false>(r8x, r8y, s, ax, ay, hm.data(), ok, g_list2[j], g_comb, BJJ_MODE_EDDSA);
    return flags;
}

uint32_t emu_verify_schnorr(size_t n, const uint8_t* pkx, const uint8_t* pky, const uint8_t* msg, const uint8_t* rx,
                            const uint8_t* ry, const uint8_t* s, uint8_t* ok, uint8_t* status) {
    emu_init();
    uint32_t flags = 0;
    ExactQueue qa = exact_queue(n), qr = exact_queue2(n);
    std::vector<uint8_t> hm(4 * 32 * (n + 1));
    for (size_t i = 0; i < n; i++)
        lane_verify_hash(rx, ry, pkx, pky, msg, s, 1, 0, nullptr, hm.data(), n, ok, i, true, qa, qr, flags, BJJ_MODE_SCHNORR, g_split, status);
    for (size_t i = 0; i < n; i++) lane_verify_ec(rx, ry, pkx, pky, hm.data(), n, ok, i, lane_table(0), lane_table(1), g_comb, BJJ_MODE_SCHNORR, host_park());
    for (uint32_t j = 0; j < g_count; j++) lane_verify_exact<true>(rx, ry, s, pkx, pky, hm.data(), ok, g_list[j], g_comb, BJJ_MODE_SCHNORR);
    for (uint32_t j = 0; j < g_count2; j++) lane_verify_exact<false>(rx, ry, s, pkx, pky, hm.data(), ok, g_list2[j], g_comb, BJJ_MODE_SCHNORR);
    return flags;
}

void emu_verify_compressed(size_t n, const uint8_t* sig64, const uint8_t* pk32, const uint8_t* msg, uint8_t* ok,
                           uint8_t* status) {
    emu_init();
    std::vector<uint8_t> d(4 * 32 * (n + 1)), hm(4 * 32 * (n + 1));
    uint8_t *dx = d.data(), *dy = dx + 32 * n, *ax = dy + 32 * n, *ay = ax + 32 * n;
    uint32_t flags = 0;
    ExactQueue q = exact_queue(n);
    ProjScratch scr = proj_scratch(2 * n);
    for (size_t i = 0; i < n; i++) lane_decompress_prepare(sig64, 2, 0, scr, i, i);
    for (size_t i = 0; i < n; i++) lane_decompress_prepare(pk32, 1, 0, scr, n + i, i);
    batch_inverse(scr, 2 * n);
    for (size_t i = 0; i < n; i++) lane_decompress_finish(sig64, 2, 0, scr, i, dx, dy, status, i, false);
    for (size_t i = 0; i < n; i++) lane_decompress_finish(pk32, 1, 0, scr, n + i, ax, ay, status, i, true);
    for (size_t i = 0; i < n; i++)
        lane_verify_hash(dx, dy, ax, ay, msg, sig64, 2, 1, status, hm.data(), n, ok, i, false, q, q, flags, BJJ_MODE_EDDSA, g_split, nullptr);
    if (g_split)
        for (size_t i = 0; i < n; i++) lane_verify_split(sig64, 2, 1, hm.data(), n, ok, i);
    for (size_t i = 0; i < n; i++) lane_verify_ec(dx, dy, ax, ay, hm.data(), n, ok, i, lane_table(0), lane_table(1), g_comb, BJJ_MODE_EDDSA, host_park());
}

// verify's half-size scalar split (split.cuh), for the invariant tests: u = v * h (mod l), v odd
void emu_set_split(int on) { g_split = on != 0; }
void emu_split(size_t n, const uint8_t* h32, const uint8_t* s32, uint8_t* u32, uint8_t* v32, uint8_t* vneg, uint8_t* w32) {
    for (size_t i = 0; i < n; i++) {
        uint32_t h[8], s[8], u[8], v[8], w[8], neg = 0;
        load_u256(h, h32, i);
        load_u256(s, s32, i);
        split_scalars(u, v, neg, h);
        split_scale_s(w, s, v);
        store_u256(u32, i, u);
        store_u256(v32, i, v);
        store_u256(w32, i, w);
        vneg[i] = (uint8_t)neg;
    }
}

}  // extern "C"
