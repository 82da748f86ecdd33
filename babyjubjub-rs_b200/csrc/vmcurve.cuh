// Twisted-Edwards arithmetic (a = -1 model, extended coordinates) over shared-memory slots: the formulas of
// curve.cuh (ext_dbl, ext_add_niels, ext_add_niels_aff, niels_from_ext) expressed as sequences of the out-of-line
// operations of vm.cuh.  A doubling is 7 calls, an addition 7; the instruction footprint of a whole Straus pass is
// a few KB of call sites plus ~10 KB of subroutines, instead of 150 KB of inlined multiplications per window.
//
// Reference items: PointProjective::add (src/lib.rs:88-131) and Point::mul_scalar (src/lib.rs:149-164) for inputs ON
// the curve, where any correct group arithmetic yields the reference's canonical result.
#pragma once
#include "lanes.cuh"
#include "vm.cuh"

#if defined(__CUDACC__) && !defined(BJJ_HOST_EMU)

namespace bjj {
namespace vm {

enum { C_R2 = 0, C_SQRT_NEG_A = 1, C_TWO_DP = 2 };

// d = a * K for a library constant K (K is the multiplicand: `a` may then be any 256-bit integer, which is what
// the conversion of caller bytes to Montgomery form needs -- see fr_to_mont)
static __device__ __noinline__ void mul_c(Slot d, Slot a, int which) {
    Fr x, c, r;
    ld(x, a);
    if (which == C_R2)
        c = fr_const(BJJ_R2);
    else if (which == C_SQRT_NEG_A)
        c = fr_const(BJJ_SQRT_NEG_A_M);
    else
        c = fr_const(BJJ_TWO_DP_M);
    fr_mul_inline(r, c, x);
    st(d, r);
}
static __device__ __noinline__ void neg(Slot d, Slot a) {
    Fr x, r;
    ld(x, a);
    fr_neg(r, x);
    st(d, r);
}
// d = 2a - b
static __device__ __noinline__ void dblsub(Slot d, Slot a, Slot b) {
    BJJ_VM_SMALL_OP_BEGIN
    Fr x, y, r;
    ld(x, a);
    ld(y, b);
    fr_add(r, x, x);
    fr_sub(r, r, y);
    st(d, r);
}
// slot -> global (two 16-byte halves `hstride` uint4 apart)
static __device__ __noinline__ void put(uint4* g, size_t hstride, Slot a) {
    Fr x;
    ld(x, a);
    stg(g, hstride, x);
}
// 0 or 1 (Montgomery form)
static __device__ __noinline__ void set01(Slot d, int one) {
    Fr r = fr_const(BJJ_ONE_M);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = one ? r.v[i] : 0u;
    st(d, r);
}

// the working set of one lane: accumulator (X, Y, Z, T) and five temporaries
struct Regs {
    Slot X, Y, Z, T, t0, t1, t2, t3, t4;
};
__device__ __forceinline__ Regs regs_at(int first_slot) {
    Regs s;
    s.X = slot(first_slot);
    s.Y = slot(first_slot + 1);
    s.Z = slot(first_slot + 2);
    s.T = slot(first_slot + 3);
    s.t0 = slot(first_slot + 4);
    s.t1 = slot(first_slot + 5);
    s.t2 = slot(first_slot + 6);
    s.t3 = slot(first_slot + 7);
    s.t4 = slot(first_slot + 8);
    return s;
}
#define BJJ_VM_REG_SLOTS 9

__device__ __forceinline__ void set_identity(const Regs& s) {
    set01(s.X, 0);
    set01(s.Y, 1);
    set01(s.Z, 1);
    set01(s.T, 0);
}

// acc = 2 acc   (dbl-2008-hwcd, a = -1: 4S + 3M, +1M for T; curve.cuh::ext_dbl)
__device__ __forceinline__ void dbl(const Regs& s, bool want_t) {
    sqr2(s.t0, s.X, s.t1, s.Y, s.Y, 0);          // xx, yy
    sqr2(s.t2, s.Z, s.t3, s.X, s.Y, 1);          // zz, (X+Y)^2
    addsub(s.t4, s.t1, s.t1, s.t0);              // H' = yy + xx, G = yy - xx
    sub(s.t3, s.t3, s.t4);                       // E = 2XY
    dblsub(s.t2, s.t2, s.t1);                    // F' = 2zz - G
    mul2(s.X, s.t3, s.t2, s.Y, s.t4, s.t1);      // X = E F', Y = H' G
    if (want_t)
        mul2(s.Z, s.t1, s.t2, s.T, s.t3, s.t4);  // Z = G F', T = E H'
    else
        mul(s.Z, s.t1, s.t2);
}

// acc += (-1)^negate * entry, the entry in GLOBAL memory as (y+x, y-x, 2d'T[, 2Z]): coordinate c at e + c * cstride,
// halves `hstride` apart.  affine: Z(entry) = 1, the fourth coordinate is not read.  (curve.cuh::ext_add_niels[_aff];
// the negation swaps the first two coordinates and the roles of F and G instead of negating 2d'T.)
// The entry's coordinates travel global -> caller registers -> free slots: the loads are issued one subroutine call
// ahead of their use (their L2/HBM latency runs under ~70 or ~1,000 instructions of arithmetic), and the products then
// go through the SAME mul2/mul as everything else.  A mul2 variant with its second factors in global memory was the
// first design; with the squaring subroutine added it pushed the hot code of the Straus loop from 23 to 29 KB, past
// the instruction cache, and the kernel fell to 79 % multiplier-pipe activity with 1.9 `no_instruction` stalls per
// issue (profiles/r2_ncu_verify_ec_summary.txt).
__device__ __forceinline__ void add_entry(const Regs& s, const uint4* e, size_t cstride, size_t hstride, bool negate, bool affine,
                                          bool want_t) {
    Fr g0, g1;
    ldg(g0, e + (negate ? cstride : 0), hstride);
    ldg(g1, e + (negate ? 0 : cstride), hstride);
    addsub(s.t0, s.t1, s.Y, s.X);                                                                   // Y+X, Y-X
    st(s.t2, g0);
    st(s.t3, g1);
    ldg(g0, e + 2 * cstride, hstride);
    if (!affine) ldg(g1, e + 3 * cstride, hstride);
    mul2(s.t0, s.t0, s.t2, s.t1, s.t1, s.t3);                                                       // A, B
    st(s.t2, g0);
    if (affine) {
        mul(s.t2, s.T, s.t2);                                                                       // C
        add(s.t3, s.Z, s.Z);                                                                        // D = 2Z
    } else {
        st(s.t4, g1);
        mul2(s.t2, s.T, s.t2, s.t3, s.Z, s.t4);                                                     // C, D
    }
    addsub(s.t4, s.t0, s.t0, s.t1);                                                                 // H = A + B, E = A - B
    addsub(negate ? s.t2 : s.t1, negate ? s.t1 : s.t2, s.t3, s.t2);                                 // G -> t1, F -> t2
    mul2(s.X, s.t0, s.t2, s.Y, s.t1, s.t4);                                                         // X = E F, Y = G H
    if (want_t)
        mul2(s.Z, s.t2, s.t1, s.T, s.t0, s.t4);                                                     // Z = F G, T = E H
    else
        mul(s.Z, s.t2, s.t1);
}

// Niels form of acc -> global entry (curve.cuh::niels_from_ext); clobbers t0..t3
__device__ __forceinline__ void store_niels(const Regs& s, uint4* e, size_t cstride, size_t hstride) {
    addsub(s.t0, s.t1, s.Y, s.X);
    mul_c(s.t2, s.T, C_TWO_DP);
    add(s.t3, s.Z, s.Z);
    put(e, hstride, s.t0);
    put(e + cstride, hstride, s.t1);
    put(e + 2 * cstride, hstride, s.t2);
    put(e + 3 * cstride, hstride, s.t3);
}

// Per-thread window table in global memory: entry j = j * P (j = 0..8) is 128 contiguous bytes (y+x | y-x | 2d'T | 2Z)
// and a thread's nine entries are contiguous.  The lanes of a warp pick different entries, so what matters for L2
// and HBM is that every 32-byte sector fetched is used: with the entry in one 128-byte line it is (the interleaved
// layout of lanes.cuh::LaneTable, which coalesces only when all lanes pick the same entry, fetched twice the bytes).
struct Table {
    uint4* base;      // entry 0 of THIS thread
    __device__ __forceinline__ uint4* entry(int j) const { return base + (size_t)j * 8; }
};
#define BJJ_VM_TABLE_CSTRIDE 2      // uint4 between coordinates of an entry
#define BJJ_VM_TABLE_HSTRIDE 1      // uint4 between the halves of a coordinate
__device__ __forceinline__ Table table_of(U128* pool, size_t thread_slot) {
    Table r;
    r.base = reinterpret_cast<uint4*>(pool) + thread_slot * BJJ_TABLE_U128_PER_LANE;
    return r;
}

// entries 0..8 = j * acc (acc is left at 8P)
__device__ __forceinline__ void table_build(const Regs& s, const Table& tb) {
    // entry 0: the identity (1, 1, 0, 2); built through the slots so that no second code path exists
    set01(s.t0, 1);
    set01(s.t2, 0);
    add(s.t3, s.t0, s.t0);
    put(tb.entry(0), BJJ_VM_TABLE_HSTRIDE, s.t0);
    put(tb.entry(0) + BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE, s.t0);
    put(tb.entry(0) + 2 * BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE, s.t2);
    put(tb.entry(0) + 3 * BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE, s.t3);
    store_niels(s, tb.entry(1), BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE);
#pragma unroll 1
    for (int j = 2; j <= 8; j++) {
        add_entry(s, tb.entry(1), BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE, false, false, true);
        store_niels(s, tb.entry(j), BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE);
    }
}

// acc += d * P for a signed digit d in [-8, 8]
__device__ __forceinline__ void add_digit(const Regs& s, const Table& tb, int d, bool want_t) {
    const int ad = d < 0 ? -d : d;
    add_entry(s, tb.entry(ad), BJJ_VM_TABLE_CSTRIDE, BJJ_VM_TABLE_HSTRIDE, d < 0, false, want_t);
}
// acc += d * 65536^w * B8 from the fixed-base table (lanes.cuh::comb_select)
__device__ __forceinline__ void add_comb(const Regs& s, const CombEntry* comb, int w, int d, bool want_t) {
    const int ad = d < 0 ? -d : d;
    const uint4* e = reinterpret_cast<const uint4*>(comb + (size_t)w * BJJ_COMB_ENTRIES + ad);
    add_entry(s, e, 2, 1, d < 0, true, want_t);
}

// caller bytes (canonical integer, 32-byte element i of `base`) -> Montgomery form in slot d
__device__ __forceinline__ void load_mont(Slot d, const uint8_t* base, size_t i) {
    Fr raw;
    load_u256(raw.v, base, i);
    st(d, raw);
    mul_c(d, d, C_R2);
}

// affine point of the ORIGINAL curve (Montgomery x, y already in s.X, s.Y) -> extended point on the a = -1 model
__device__ __forceinline__ void from_affine(const Regs& s) {
    mul_c(s.X, s.X, C_SQRT_NEG_A);
    set01(s.Z, 1);
    mul(s.T, s.X, s.Y);
}

}  // namespace vm
}  // namespace bjj

#endif
