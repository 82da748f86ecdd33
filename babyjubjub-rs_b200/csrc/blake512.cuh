// BLAKE-512 (the original SHA-3 finalist, NOT BLAKE2) for short, single-block inputs.
// Replaces blake-hash 0.4.0 `Blake512::digest` behind the reference's blh() (src/lib.rs:226-237),
// whose only inputs are the 32-byte key (src/lib.rs:291,316) and the 64-byte h[32..64] || msg32
// (src/lib.rs:326) -- both fit one 128-byte block.
#pragma once
#include <stdint.h>
#include "fr.cuh"

namespace bjj {

BJJ_HD uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

BJJ_HD uint64_t blake_c(int i) {
    switch (i) {
        case 0: return 0x243F6A8885A308D3ull;
        case 1: return 0x13198A2E03707344ull;
        case 2: return 0xA4093822299F31D0ull;
        case 3: return 0x082EFA98EC4E6C89ull;
        case 4: return 0x452821E638D01377ull;
        case 5: return 0xBE5466CF34E90C6Cull;
        case 6: return 0xC0AC29B7C97C50DDull;
        case 7: return 0x3F84D5B5B5470917ull;
        case 8: return 0x9216D5D98979FB1Bull;
        case 9: return 0xD1310BA698DFB5ACull;
        case 10: return 0x2FFD72DBD01ADFB7ull;
        case 11: return 0xB8E1AFED6A267E96ull;
        case 12: return 0xBA7C9045F12C7F99ull;
        case 13: return 0x24A19947B3916CF7ull;
        case 14: return 0x0801F2E2858EFC16ull;
        default: return 0x636920D871574E69ull;
    }
}

// sigma rows packed 4 bits per entry (entry k in bits [4k, 4k+4))
BJJ_HD uint64_t blake_sigma_row(int r) {
    switch (r) {
        case 0: return 0xFEDCBA9876543210ull;
        case 1: return 0x357B20C16DF984AEull;
        case 2: return 0x491763EADF250C8Bull;
        case 3: return 0x8F04A562EBCD1397ull;
        case 4: return 0xD386CB1EFA427509ull;
        case 5: return 0x91EF57D438B0A6C2ull;
        case 6: return 0xB8293670A4DEF15Cull;
        case 7: return 0xA2684F05931CE7BDull;
        case 8: return 0x5A417D2C803B9EF6ull;
        default: return 0x0DC3E9BF5167482Aull;
    }
}

// One compression of a single padded block `m` (16 big-endian words already decoded) with bit
// counter t (< 2^64).  h is updated in place.
BJJ_HD void blake512_compress(uint64_t* h, const uint64_t* m, uint64_t t) {
    uint64_t v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = h[i];
    v[8] = blake_c(0);
    v[9] = blake_c(1);
    v[10] = blake_c(2);
    v[11] = blake_c(3);
    v[12] = blake_c(4) ^ t;
    v[13] = blake_c(5) ^ t;
    v[14] = blake_c(6);
    v[15] = blake_c(7);
#define BJJ_BLAKE_G(a, b, c, d, i)                                          \
    {                                                                       \
        int s0 = (int)((sig >> (8 * (i))) & 15), s1 = (int)((sig >> (8 * (i) + 4)) & 15); \
        v[a] = v[a] + v[b] + (m[s0] ^ blake_c(s1));                         \
        v[d] = ror64(v[d] ^ v[a], 32);                                      \
        v[c] = v[c] + v[d];                                                 \
        v[b] = ror64(v[b] ^ v[c], 25);                                      \
        v[a] = v[a] + v[b] + (m[s1] ^ blake_c(s0));                         \
        v[d] = ror64(v[d] ^ v[a], 16);                                      \
        v[c] = v[c] + v[d];                                                 \
        v[b] = ror64(v[b] ^ v[c], 11);                                      \
    }
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const uint64_t sig = blake_sigma_row(r % 10);
        BJJ_BLAKE_G(0, 4, 8, 12, 0)
        BJJ_BLAKE_G(1, 5, 9, 13, 1)
        BJJ_BLAKE_G(2, 6, 10, 14, 2)
        BJJ_BLAKE_G(3, 7, 11, 15, 3)
        BJJ_BLAKE_G(0, 5, 10, 15, 4)
        BJJ_BLAKE_G(1, 6, 11, 12, 5)
        BJJ_BLAKE_G(2, 7, 8, 13, 6)
        BJJ_BLAKE_G(3, 4, 9, 14, 7)
    }
#undef BJJ_BLAKE_G
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

BJJ_HD uint64_t bswap64(uint64_t x) {
    x = ((x & 0x00FF00FF00FF00FFull) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFull);
    x = ((x & 0x0000FFFF0000FFFFull) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFull);
    return (x << 32) | (x >> 32);
}

// digest of a message given as NW big-endian-decoded 64-bit words (NW*8 bytes, NW <= 13)
template <int NW>
BJJ_HD void blake512_short_be(uint64_t* h, const uint64_t* be) {
    uint64_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) m[i] = be[i];
    m[NW] = 0x8000000000000000ull;
    m[13] |= 1ull;                 // byte 111 = 0x01 (BLAKE-512, not BLAKE-384)
    m[15] = (uint64_t)NW * 64;     // message length in bits
    h[0] = 0x6A09E667F3BCC908ull;
    h[1] = 0xBB67AE8584CAA73Bull;
    h[2] = 0x3C6EF372FE94F82Bull;
    h[3] = 0xA54FF53A5F1D36F1ull;
    h[4] = 0x510E527FADE682D1ull;
    h[5] = 0x9B05688C2B3E6C1Full;
    h[6] = 0x1F83D9ABFB41BD6Bull;
    h[7] = 0x5BE0CD19137E2179ull;
    blake512_compress(h, m, (uint64_t)NW * 64);
}

// digest of a message of NW little-endian-loaded 64-bit words (NW*8 bytes, NW <= 13): `le[k]` holds
// message bytes 8k..8k+7 as read by a little-endian load.  Output: 8 big-endian state words.
template <int NW>
BJJ_HD void blake512_short(uint64_t* h, const uint64_t* le) {
    uint64_t be[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) be[i] = bswap64(le[i]);
    blake512_short_be<NW>(h, be);
}

// RFC 8032 pruning + >> 3 of digest bytes 0..31 (= state words h[0..3], big-endian)
BJJ_HD void scalar_key_from_digest(uint32_t* out, const uint64_t* h) {
    uint64_t l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) l[i] = bswap64(h[i]);
    l[0] &= ~0x07ull;                          // h[0] &= 0xF8
    l[3] &= 0x7FFFFFFFFFFFFFFFull;             // h[31] &= 0x7F
    l[3] |= 0x4000000000000000ull;             // h[31] |= 0x40
    uint64_t s[4];
    s[0] = (l[0] >> 3) | (l[1] << 61);
    s[1] = (l[1] >> 3) | (l[2] << 61);
    s[2] = (l[2] >> 3) | (l[3] << 61);
    s[3] = l[3] >> 3;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out[2 * i] = (uint32_t)s[i];
        out[2 * i + 1] = (uint32_t)(s[i] >> 32);
    }
}

// PrivateKey::scalar_key (src/lib.rs:284-302): BLAKE-512(key)[..32], RFC 8032 pruning, >> 3.
// key: 8 u32 little-endian-loaded words.  out: 256-bit scalar, 8 u32 limbs.
BJJ_HD void scalar_key_from_key(uint32_t* out, const uint32_t* key) {
    uint64_t le[4], h[8];
#pragma unroll
    for (int i = 0; i < 4; i++) le[i] = (uint64_t)key[2 * i] | ((uint64_t)key[2 * i + 1] << 32);
    blake512_short<4>(h, le);
    scalar_key_from_digest(out, h);
}

}  // namespace bjj
