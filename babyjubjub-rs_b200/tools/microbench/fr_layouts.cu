// Which layout should the Montgomery multiplier of the EC kernels use?  Measured, not assumed
// (BASELINE.json north_star: "one thread or a warp-cooperative layout depending on measured register pressure").
//
//   footprint   one thread per element, N inlined multiplications in a straight-line loop body (3 KB of SASS each),
//               2 CTAs x 128 threads per SM like k_verify_ec, warps staggered: throughput against body size shows the
//               instruction-cache cliff that bounded round 1's Straus kernel.
//   vm_mul      one thread per element, operands in shared memory, ONE out-of-line multiplier (csrc/vm.cuh);
//               vm_mul2 = two multiplications per call; vm_dbl = the real doubling formula (4 sqr, 4 mul, 6 add/sub).
//   coop<T>     T = 2 or 4 threads per element (4 or 2 limbs each): b-limb broadcast, per-thread partial chains,
//               limb hand-over and carry resolution by warp shuffles.
//
// Every variant is validated on the device against fr_mul_inline / ext_dbl before it is timed.  One JSON object per
// line.  Timed with CUDA events after a warm-up.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../csrc/curve.cuh"
#include "../../csrc/fr.cuh"
#include "../../csrc/vm.cuh"

using namespace bjj;

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e = (x);                                                                        \
        if (e != cudaSuccess) {                                                                     \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                                \
        }                                                                                           \
    } while (0)

__device__ __forceinline__ void seed_fr(Fr& x, uint32_t s) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s = s * 1664525u + 1013904223u;
        x.v[i] = s;
    }
    x.v[7] &= 0x1fffffffu;      // < 2^253 < Q
}

// ---------------------------------------------------------------------------------------------------
// 1. footprint sweep
// ---------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 2) k_footprint(uint32_t* out, int iters, uint32_t seed, int stagger) {
    Fr x0, x1, y;
    seed_fr(x0, seed + threadIdx.x);
    seed_fr(x1, seed * 3 + threadIdx.x);
    seed_fr(y, seed * 7 + blockIdx.x);
    // warps of an SM start out of step, as they end up in a long kernel
    const int lead = ((threadIdx.x >> 5) + 4 * (blockIdx.x & 1)) * stagger;
#pragma unroll 1
    for (int i = 0; i < lead; i++) fr_mul(x0, x0, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            fr_mul(x0, x0, y);
            fr_mul(x1, x1, y);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x0.v[i] ^ x1.v[i];
    if (s == 0x1234567) out[0] = s;
}

// ---------------------------------------------------------------------------------------------------
// 2. shared-memory operands, out-of-line multiplier
// ---------------------------------------------------------------------------------------------------
// MODE 0: vm::mul, 4 dependent streams;  1: vm::mul2;  2: the doubling formula (dbl-2008-hwcd, a = -1) on slots
template <int MODE>
__global__ void __launch_bounds__(128) k_vm(uint32_t* out, int iters, uint32_t seed) {
    using namespace vm;
    const Slot X = slot(0), Y = slot(1), Z = slot(2), T = slot(3), t0 = slot(4), t1 = slot(5), t2 = slot(6), t3 = slot(7),
               t4 = slot(8);
    Fr v;
    for (int s = 0; s < 9; s++) {
        seed_fr(v, seed + 977 * s + threadIdx.x + 131 * blockIdx.x);
        st(slot(s), v);
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            mul(X, X, t0);
            mul(Y, Y, t0);
            mul(Z, Z, t0);
            mul(T, T, t0);
        } else if (MODE == 1) {
            mul2(X, X, t0, Y, Y, t0);
            mul2(Z, Z, t0, T, T, t0);
        } else {
            mul2(t0, X, X, t1, Y, Y);
            add(t3, X, Y);
            mul2(t2, Z, Z, t3, t3, t3);
            addsub(t4, t1, t1, t0);      // h = yy + xx, g = yy - xx
            sub(t3, t3, t4);             // e = s - h
            add(t2, t2, t2);
            sub(t2, t2, t1);             // f = 2zz - g
            mul2(X, t3, t2, Y, t4, t1);
            mul2(Z, t1, t2, T, t3, t4);
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < 4; k++) {
        ld(v, slot(k));
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= v.v[i];
    }
    if (s == 0x1234567) out[0] = s;
}

// the doubling on slots against ext_dbl<true> in registers
__global__ void __launch_bounds__(128) k_vm_check(uint32_t* bad, uint32_t seed) {
    using namespace vm;
    const Slot X = slot(0), Y = slot(1), Z = slot(2), T = slot(3), t0 = slot(4), t1 = slot(5), t2 = slot(6), t3 = slot(7),
               t4 = slot(8);
    PointExt p, r;
    seed_fr(p.X, seed + threadIdx.x);
    seed_fr(p.Y, seed * 5 + threadIdx.x);
    seed_fr(p.Z, seed * 9 + threadIdx.x);
    seed_fr(p.T, seed * 11 + threadIdx.x);
    st(X, p.X);
    st(Y, p.Y);
    st(Z, p.Z);
    st(T, p.T);
    for (int it = 0; it < 3; it++) {
        ext_dbl<true>(r, p);
        p = r;
        mul2(t0, X, X, t1, Y, Y);
        add(t3, X, Y);
        mul2(t2, Z, Z, t3, t3, t3);
        addsub(t4, t1, t1, t0);
        sub(t3, t3, t4);
        add(t2, t2, t2);
        sub(t2, t2, t1);
        mul2(X, t3, t2, Y, t4, t1);
        mul2(Z, t1, t2, T, t3, t4);
        mul(t0, X, Y);       // exercise the single multiplier too
        fr_mul(r.X, p.X, p.Y);
        Fr c;
        ld(c, t0);
        if (!u256_eq(c.v, r.X.v)) atomicAdd(bad, 1u);
    }
    Fr a, b, c, d;
    ld(a, X);
    ld(b, Y);
    ld(c, Z);
    ld(d, T);
    if (!u256_eq(a.v, p.X.v) || !u256_eq(b.v, p.Y.v) || !u256_eq(c.v, p.Z.v) || !u256_eq(d.v, p.T.v)) atomicAdd(bad, 1u);
}

// ---------------------------------------------------------------------------------------------------
// 3. warp-cooperative multiplier: T threads per element, L = 8 / T limbs per thread
// ---------------------------------------------------------------------------------------------------
// acc[0 .. L+1] += {a[0..L-1]} * b.  Even-indexed products form one carry chain, odd-indexed ones a second chain one
// column up (the same trick as fr.cuh::mac4), tops rippling into acc[L], acc[L+1].
template <int L>
__device__ __forceinline__ void coop_mac(uint32_t* acc, const uint32_t* a, uint32_t b);
template <>
__device__ __forceinline__ void coop_mac<2>(uint32_t* acc, const uint32_t* a, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\tmadc.hi.cc.u32 %2, %5, %6, %2;\n\taddc.u32 %3, %3, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
        : "r"(a[0]), "r"(a[1]), "r"(b));
}
template <>
__device__ __forceinline__ void coop_mac<4>(uint32_t* acc, const uint32_t* a, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %6, %10, %0;\n\tmadc.hi.cc.u32 %1, %6, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\taddc.u32 %5, %5, 0;\n\t"
        "mad.lo.cc.u32 %1, %7, %10, %1;\n\tmadc.hi.cc.u32 %2, %7, %10, %2;\n\t"
        "madc.lo.cc.u32 %3, %9, %10, %3;\n\tmadc.hi.cc.u32 %4, %9, %10, %4;\n\taddc.u32 %5, %5, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
}

// r = a * b / 2^256 mod Q on the distributed layout (thread t of the group holds limbs tL .. tL+L-1; q = its limbs of Q).
// The same interleaved CIOS as fr_mul_inline, so the result is the same integer in [0, 2Q).
template <int T>
__device__ __forceinline__ void coop_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* q) {
    constexpr int L = 8 / T;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned g0 = lane & ~(unsigned)(T - 1);
    const bool top = (lane & (T - 1)) == (T - 1);
    uint32_t acc[L + 2];
#pragma unroll
    for (int k = 0; k < L + 2; k++) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t bi = __shfl_sync(0xffffffffu, b[i % L], g0 + i / L);
        coop_mac<L>(acc, a, bi);
        const uint32_t m = __shfl_sync(0xffffffffu, acc[0] * BJJ_NINV32, g0);
        coop_mac<L>(acc, q, m);
        // divide by 2^32: every thread hands its lowest limb to the thread below (the group's lowest limb is 0 now)
        uint32_t in = __shfl_down_sync(0xffffffffu, acc[0], 1);
        if (top) in = 0;
#pragma unroll
        for (int k = 0; k <= L; k++) acc[k] = acc[k + 1];
        acc[L + 1] = 0;
        // `in` belongs to column L-1 of the shifted window
        if (L == 2)
            asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(acc[1]), "+r"(acc[2]) : "r"(in));
        else
            asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(acc[3]), "+r"(acc[4]) : "r"(in));
    }
    // carry resolution: the overflow limb acc[L] of thread t belongs to limb 0 of thread t+1
#pragma unroll 1
    while (__any_sync(0xffffffffu, acc[L] != 0)) {
        uint32_t ov = __shfl_up_sync(0xffffffffu, acc[L], 1);
        if ((lane & (T - 1)) == 0) ov = 0;
        acc[L] = 0;
        if (L == 2)
            asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 %2, %2, 0;" : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(ov));
        else
            asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;"
                : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
                : "r"(ov));
    }
#pragma unroll
    for (int k = 0; k < L; k++) r[k] = acc[k];
}

template <int T>
__device__ __forceinline__ void coop_q(uint32_t* q) {
    constexpr int L = 8 / T;
    const uint32_t Q[8] = BJJ_LIMBS8(BJJ_Q);
    const unsigned t = threadIdx.x & (T - 1);
#pragma unroll
    for (int k = 0; k < L; k++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < T; j++) v = (t == (unsigned)j) ? Q[j * L + k] : v;
        q[k] = v;
    }
}

template <int T, int ILP>
__global__ void __launch_bounds__(128) k_coop(uint32_t* out, int iters, uint32_t seed) {
    constexpr int L = 8 / T;
    uint32_t q[L], x[ILP][L], y[ILP][L];
    coop_q<T>(q);
    const unsigned t = threadIdx.x & (T - 1);
    const unsigned elem = (blockIdx.x * blockDim.x + threadIdx.x) / T;
#pragma unroll
    for (int s = 0; s < ILP; s++) {
        Fr fx, fy;
        seed_fr(fx, seed + elem * 31 + s);
        seed_fr(fy, seed * 5 + elem * 17 + s);
#pragma unroll
        for (int k = 0; k < L; k++) {
            uint32_t vx = 0, vy = 0;
#pragma unroll
            for (int j = 0; j < T; j++) {
                vx = (t == (unsigned)j) ? fx.v[j * L + k] : vx;
                vy = (t == (unsigned)j) ? fy.v[j * L + k] : vy;
            }
            x[s][k] = vx;
            y[s][k] = vy;
        }
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < ILP; s++) coop_mul<T>(x[s], x[s], y[s], q);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < L; i++) s ^= x[k][i];
    if (s == 0x1234567) out[0] = s;
}

// cooperative product against fr_mul_inline on the same operands (every thread of a group recomputes the reference)
template <int T>
__global__ void __launch_bounds__(128) k_coop_check(uint32_t* bad, uint32_t seed, int rounds) {
    constexpr int L = 8 / T;
    uint32_t q[L], x[L], y[L];
    coop_q<T>(q);
    const unsigned t = threadIdx.x & (T - 1);
    const unsigned elem = (blockIdx.x * blockDim.x + threadIdx.x) / T;
    Fr fx, fy;
    seed_fr(fx, seed + elem * 31);
    seed_fr(fy, seed * 5 + elem * 17);
    if ((elem & 7) == 1) {      // stress the carries: operands near 2Q - 1 (the lazy domain's upper end)
        const uint32_t twoq[8] = BJJ_LIMBS8(BJJ_2Q);
#pragma unroll
        for (int i = 0; i < 8; i++) fx.v[i] = fy.v[i] = twoq[i];
        fx.v[0] -= 1 + (elem >> 3);
        fy.v[0] -= 2;
    }
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int k = 0; k < L; k++) {
            uint32_t vx = 0, vy = 0;
#pragma unroll
            for (int j = 0; j < T; j++) {
                vx = (t == (unsigned)j) ? fx.v[j * L + k] : vx;
                vy = (t == (unsigned)j) ? fy.v[j * L + k] : vy;
            }
            x[k] = vx;
            y[k] = vy;
        }
        coop_mul<T>(x, x, y, q);
        Fr ref;
        fr_mul_inline(ref, fx, fy);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < L; k++) {
            uint32_t want = 0;
#pragma unroll
            for (int j = 0; j < T; j++) want = (t == (unsigned)j) ? ref.v[j * L + k] : want;
            ok = ok && (want == x[k]);
        }
        if (!ok) atomicAdd(bad, 1u);
        fy = fx;
        fx = ref;
    }
}

// ---------------------------------------------------------------------------------------------------
template <class K>
static float time_kernel(K launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static int g_sms = 0;
static uint32_t* g_out = nullptr;

template <int N>
static void run_footprint() {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_footprint<N>));
    for (int stagger : {0, 5}) {
        const int grid = g_sms * 2, iters = 12288 / N;
        float ms = time_kernel([&]() { k_footprint<N><<<grid, 128>>>(g_out, iters, 777u, stagger); }, 3);
        // the staggered lead-in is part of the timed work
        double fm = (double)grid * 128 * iters * N;
        for (int b = 0; b < grid; b++)
            for (int w = 0; w < 4; w++) fm += 32.0 * (w + 4 * (b & 1)) * stagger;
        printf("{\"bench\": \"footprint\", \"inlined_fmul\": %d, \"body_kb\": %.1f, \"stagger\": %d, \"regs\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n", N,
               N * 3.0, stagger, fa.numRegs, ms, fm / ms / 1e6);
        fflush(stdout);
    }
}

template <int MODE>
static void run_vm(const char* name, double fmul_per_iter) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_vm<MODE>));
    CK(cudaFuncSetAttribute(k_vm<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t need = vm::slot_bytes(9);
    for (int per_sm : {1, 2, 3, 4, 5, 6}) {
        // dynamic shared memory sized so that exactly per_sm CTAs fit on an SM
        size_t smem = (size_t)(227 * 1024) / per_sm - 1024;
        if (smem > 200 * 1024) smem = 200 * 1024;
        if (smem < need) continue;
        if ((long)fa.numRegs * 128 * per_sm > 65536) continue;
        int resident = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_vm<MODE>, 128, smem));
        const int grid = g_sms * per_sm, iters = 1024;
        float ms = time_kernel([&]() { k_vm<MODE><<<grid, 128, smem>>>(g_out, iters, 777u); }, 3);
        double fm = (double)grid * 128 * iters * fmul_per_iter;
        printf("{\"bench\": \"%s\", \"ctas_per_sm\": %d, \"resident\": %d, \"regs\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n", name, per_sm, resident,
               fa.numRegs, ms, fm / ms / 1e6);
        fflush(stdout);
    }
}

template <int T, int ILP>
static void run_coop() {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_coop<T, ILP>));
    for (int per_sm : {2, 4, 8, 12, 16}) {
        if ((long)fa.numRegs * 128 * per_sm > 65536) continue;
        const int grid = g_sms * per_sm, iters = 2048;
        float ms = time_kernel([&]() { k_coop<T, ILP><<<grid, 128>>>(g_out, iters, 777u); }, 3);
        double fm = (double)grid * 128 / T * iters * ILP;
        printf("{\"bench\": \"coop\", \"threads_per_element\": %d, \"ilp\": %d, \"ctas_per_sm\": %d, \"regs\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n", T,
               ILP, per_sm, fa.numRegs, ms, fm / ms / 1e6);
        fflush(stdout);
    }
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    CK(cudaMalloc(&g_out, 64));
    CK(cudaMemset(g_out, 0, 64));
    char uuid[40];
    for (int i = 0; i < 16; i++) sprintf(uuid + 2 * i, "%02x", (unsigned char)p.uuid.bytes[i]);
    printf("{\"bench\": \"device\", \"name\": \"%s\", \"sms\": %d, \"uuid\": \"%s\"}\n", p.name, g_sms, uuid);

    // correctness first
    {
        uint32_t* bad = g_out + 4;
        uint32_t h = 0;
        CK(cudaFuncSetAttribute(k_vm_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k_vm_check<<<8, 128, vm::slot_bytes(9)>>>(bad, 4242u);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
        printf("{\"check\": \"vm_dbl_vs_ext_dbl\", \"mismatches\": %u}\n", h);
        CK(cudaMemset(bad, 0, 4));
        k_coop_check<2><<<64, 128>>>(bad, 99u, 64);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
        printf("{\"check\": \"coop2_vs_fr_mul\", \"mismatches\": %u}\n", h);
        CK(cudaMemset(bad, 0, 4));
        k_coop_check<4><<<64, 128>>>(bad, 99u, 64);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
        printf("{\"check\": \"coop4_vs_fr_mul\", \"mismatches\": %u}\n", h);
        fflush(stdout);
    }

    run_footprint<4>();
    run_footprint<8>();
    run_footprint<10>();
    run_footprint<12>();
    run_footprint<16>();
    run_footprint<24>();
    run_footprint<48>();

    run_vm<0>("vm_mul", 4);
    run_vm<1>("vm_mul2", 4);
    run_vm<2>("vm_dbl", 8);

    run_coop<2, 1>();
    run_coop<2, 2>();
    run_coop<4, 1>();
    run_coop<4, 2>();
    CK(cudaFree(g_out));
    return 0;
}
