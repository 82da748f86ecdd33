// Instruction-throughput probe for the integer / fp64 pipes of sm_100a (B200).
// Every test is 8 independent dependent-chains per thread (so nothing is loop-invariant and ptxas
// cannot hoist or strength-reduce), 256 threads x 8 CTAs per SM (16 warps per SMSP: latency hidden).
// Reports thread-ops per clock per SM.  The SASS that ptxas actually emits for each PTX form is what
// is being measured -- inspect it with `cuobjdump -sass bin/pipe_probe`.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

#define NCH 8
#define UNROLL 8

#define PROBE_KERNEL(NAME, DECL, INIT, BODY, FOLD)                         \
    __global__ void NAME(uint32_t* out, int iters, uint32_t seed) {        \
        uint32_t b = seed * 2654435761u + threadIdx.x * 40503u + 1u;       \
        DECL;                                                              \
        _Pragma("unroll") for (int k = 0; k < NCH; k++) { INIT; }          \
        for (int i = 0; i < iters; i++) {                                  \
            _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {           \
                _Pragma("unroll") for (int k = 0; k < NCH; k++) { BODY; }  \
            }                                                              \
        }                                                                  \
        uint32_t s = 0;                                                    \
        _Pragma("unroll") for (int k = 0; k < NCH; k++) { FOLD; }          \
        if (s == 0x12345678u) out[0] = s;                                  \
    }

// 1. mul.wide.u32 (IMAD.WIDE.U32 with RZ addend), chain through the low word
PROBE_KERNEL(k_mul_wide, unsigned long long p[NCH], p[k] = seed + k,
             asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p[k]) : "r"((uint32_t)p[k] | 3u), "r"((uint32_t)(p[k] >> 32) | b)),
             s ^= (uint32_t)p[k] ^ (uint32_t)(p[k] >> 32))
// 2. mad.wide.u32 with 64-bit accumulate
PROBE_KERNEL(k_mad_wide, unsigned long long p[NCH], p[k] = seed + k,
             asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(p[k]) : "r"((uint32_t)p[k]), "r"(b)),
             s ^= (uint32_t)p[k] ^ (uint32_t)(p[k] >> 32))
// 3. mad.lo.u32
PROBE_KERNEL(k_mad_lo, uint32_t p[NCH], p[k] = seed + k,
             asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(p[k]) : "r"(b)), s ^= p[k])
// 4. mad.hi.u32
PROBE_KERNEL(k_mad_hi, uint32_t p[NCH], p[k] = seed + k * 0x01010101u,
             asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(p[k]) : "r"(b)), s ^= p[k])
// 5. carry pair: mad.lo.cc + madc.hi (fused by ptxas into one IMAD.WIDE with carry-out) + addc consumer
PROBE_KERNEL(k_mad_cc, uint32_t lo[NCH]; uint32_t hi[NCH]; uint32_t c[NCH], lo[k] = seed + k; hi[k] = k; c[k] = 0,
             asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                          : "+r"(lo[k]), "+r"(hi[k]), "+r"(c[k]) : "r"(lo[k] | 1u), "r"(b)),
             s ^= lo[k] ^ hi[k] ^ c[k])
// 6. two chained carry pairs (IMAD.WIDE carry-out then IMAD.WIDE.X carry-in)
PROBE_KERNEL(k_mad_cc2, uint32_t w[NCH][4], w[k][0] = seed + k; w[k][1] = k; w[k][2] = 3 * k; w[k][3] = 7,
             asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\t"
                          "madc.lo.cc.u32 %2, %6, %5, %2;\n\tmadc.hi.u32 %3, %6, %5, %3;"
                          : "+r"(w[k][0]), "+r"(w[k][1]), "+r"(w[k][2]), "+r"(w[k][3])
                          : "r"(w[k][0] | 1u), "r"(b), "r"(w[k][2] | 1u)),
             s ^= w[k][0] ^ w[k][1] ^ w[k][2] ^ w[k][3])
// 7. fp64 fma
PROBE_KERNEL(k_dfma, double p[NCH]; double bd = 1.0 + 1e-9 * (b & 1023); double cd = 1e-12, p[k] = 1.0 + k * 1e-6,
             asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(p[k]) : "d"(bd), "d"(cd)),
             s ^= (uint32_t)__double2loint(p[k]) ^ (uint32_t)__double2hiint(p[k]))
// 8. fp32 fma
PROBE_KERNEL(k_ffma, float p[NCH]; float bf = 1.0f + 1e-7f * (b & 1023); float cf = 1e-9f, p[k] = 1.0f + k * 1e-3f,
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(p[k]) : "f"(bf), "f"(cf)), s ^= __float_as_uint(p[k]))
// 9. 3-input add (IADD3)
PROBE_KERNEL(k_iadd3, uint32_t p[NCH]; uint32_t q[NCH], p[k] = seed + k; q[k] = b + k,
             asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(p[k]) : "r"(q[k]), "r"(b)),
             s ^= p[k])
// 10. 64-bit add (IADD3 + IADD3.X pair)
PROBE_KERNEL(k_add64, unsigned long long p[NCH]; unsigned long long q = ((unsigned long long)b << 32) | seed, p[k] = seed + k,
             asm volatile("add.u64 %0, %0, %1;" : "+l"(p[k]) : "l"(q)),
             s ^= (uint32_t)p[k] ^ (uint32_t)(p[k] >> 32))
// 11. funnel shift (SHF)
PROBE_KERNEL(k_shf, uint32_t p[NCH], p[k] = seed + k,
             asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(p[k]) : "r"(b)), s ^= p[k])
// 12. lop3
PROBE_KERNEL(k_lop3, uint32_t p[NCH], p[k] = seed + k,
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(p[k]) : "r"(b), "r"(seed)), s ^= p[k])
// 13. mixed: mul.wide (RZ) + 64-bit 3-input accumulate on the alu pipe (the shape ptxas picks for radix-2^29)
PROBE_KERNEL(k_mulwide_add, unsigned long long p[NCH]; unsigned long long acc[NCH], p[k] = seed + k; acc[k] = k,
             asm volatile("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tadd.u64 %1, %1, t;\n\tmov.b64 %0, %1;\n\t}"
                          : "+l"(p[k]), "+l"(acc[k]) : "r"((uint32_t)acc[k]), "r"(b)),
             s ^= (uint32_t)acc[k] ^ (uint32_t)(acc[k] >> 32))

template <class F>
static double run(F launch, int sms, int grid, int threads, int iters, double ops_per_iter, int clock_khz) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double ops = (double)grid * threads * iters * ops_per_iter;
    double per_clk_sm = ops / (best * 1e-3) / ((double)clock_khz * 1e3) / sms;
    return per_clk_sm;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    CK(cudaMalloc(&out, 64));
    const int threads = 256, per_sm = 8, grid = sms * per_sm, iters = 512;
    const double per_iter = (double)NCH * UNROLL;
    printf("{\"probe\": \"device\", \"name\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, sms, clock_khz / 1e3);
#define RUN(K, LABEL, MULT)                                                                                          \
    {                                                                                                                \
        double v = run([&]() { K<<<grid, threads>>>(out, iters, 99u); }, sms, grid, threads, iters, per_iter * (MULT), clock_khz); \
        printf("{\"probe\": \"%s\", \"thread_ops_per_clk_per_sm\": %.2f}\n", LABEL, v);                             \
    }
    RUN(k_mul_wide, "mul.wide.u32 (IMAD.WIDE, RZ addend)", 1)
    RUN(k_mad_wide, "mad.wide.u32 (64-bit accumulate)", 1)
    RUN(k_mad_lo, "mad.lo.u32 (IMAD)", 1)
    RUN(k_mad_hi, "mad.hi.u32 (IMAD.HI)", 1)
    RUN(k_mad_cc, "mad.lo.cc+madc.hi.cc+addc (1 fused wide MAC with carry-out)", 1)
    RUN(k_mad_cc2, "2 chained fused wide MACs (carry-out, carry-in)", 2)
    RUN(k_dfma, "fma.rn.f64 (DFMA)", 1)
    RUN(k_ffma, "fma.rn.f32 (FFMA)", 1)
    RUN(k_iadd3, "2 x add.u32 (IADD3)", 1)
    RUN(k_add64, "add.u64 (IADD3 + IADD3.X)", 1)
    RUN(k_shf, "shf.r.wrap.b32 (SHF)", 1)
    RUN(k_lop3, "lop3.b32 (LOP3)", 1)
    RUN(k_mulwide_add, "mul.wide + add.u64 (MAC split over fma and alu pipes)", 1)
    CK(cudaFree(out));
    return 0;
}
