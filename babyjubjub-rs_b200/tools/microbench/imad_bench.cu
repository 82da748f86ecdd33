// Microbenchmarks that MEASURE the roofline denominator this path is bound by: the integer-multiply
// (fma-pipe) issue rate.  MEASURED_PEAKS.json has only HBM and bf16 peaks.
//
//   imad_wide   : dependency-free IMAD.WIDE.U32 accumulations (8 independent 64-bit accumulators per
//                 thread) -> thread-IMAD/s, to compare with the model 148 SMs x 64 lanes x f_SM.
//   imad_chain  : the same count of IMAD.WIDE.U32.X in carry chains of 8 (the shape fr_mul uses)
//   fr_mul<ILP> : ILP independent Montgomery multiplication streams per thread -> fmul/s and the
//                 IMAD/s they imply (136 per fmul), at several CTA sizes / occupancies.
//
// Prints one JSON object per line.  Timed with CUDA events on the launching stream after a warm-up.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../csrc/fr.cuh"
#include "fr29_proto.cuh"
#include "fr52_proto.cuh"
#include "fr_sqr_proto.cuh"

using namespace bjj;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__global__ void k_imad_wide(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    unsigned long long acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567) out[0] = (uint32_t)s;
}

__global__ void k_imad_chain(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 7, a2 = a0 * 11, a3 = a0 * 13, b = seed * 3 + blockIdx.x;
    uint32_t c[2][10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[0][k] = c[1][k] = k;
    for (int i = 0; i < iters; i++) {
        mac4<false, 1>(c[0], a0, a1, a2, a3, b);
        mac4<false, 1>(c[1], a1, a2, a3, a0, b);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 10; k++) s ^= c[0][k] ^ c[1][k];
    if (s == 0x1234567) out[0] = s;
}

template <int ILP, bool SQR>
__global__ void k_fr_mul(uint32_t* out, int iters, uint32_t seed) {
    Fr x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[k].v[i] = seed + threadIdx.x * 977 + i * 131 + k;
            y[k].v[i] = seed * 5 + blockIdx.x * 31 + i * 17 + k;
        }
        x[k].v[7] &= 0x1fffffff;
        y[k].v[7] &= 0x1fffffff;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (SQR)
                fr_sqr_dedicated(x[k], x[k]);
            else
                fr_mul(x[k], x[k], y[k]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= x[k].v[i];
    if (s == 0x1234567) out[0] = s;
}

template <int ILP, bool SQR>
__global__ void k_fr29(uint32_t* out, int iters, uint32_t seed) {
    bjj29::Fr29 x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            x[k].v[i] = (seed + threadIdx.x * 977 + i * 131 + k) & bjj29::M29;
            y[k].v[i] = (seed * 5 + blockIdx.x * 31 + i * 17 + k) & bjj29::M29;
        }
        x[k].v[8] &= 0xffffff;
        y[k].v[8] &= 0xffffff;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (SQR)
                bjj29::sqr(x[k], x[k]);
            else
                bjj29::mul(x[k], x[k], y[k]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) s ^= x[k].v[i];
    if (s == 0x1234567) out[0] = s;
}

template <int ILP>
__global__ void k_fr52(uint32_t* out, int iters, uint32_t seed) {
    bjj52::Fr52 x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
#pragma unroll
        for (int i = 0; i < 5; i++) {
            x[k].v[i] = (double)(((uint64_t)(seed + threadIdx.x * 977 + i * 131 + k) * 0x9E3779B97F4A7C15ull) >> 12);
            y[k].v[i] = (double)(((uint64_t)(seed * 5 + blockIdx.x * 31 + i * 17 + k) * 0xC2B2AE3D27D4EB4Full) >> 12);
        }
        x[k].v[4] = (double)((uint64_t)x[k].v[4] >> 8);
        y[k].v[4] = (double)((uint64_t)y[k].v[4] >> 8);
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) bjj52::mul(x[k], x[k], y[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 5; i++) s += x[k].v[i];
    if (s == 1234567.0) out[0] = 1;
}

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
    return best;
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    CK(cudaMalloc(&out, 64));
    const double model_peak = (double)sms * 64.0 * (double)clock_khz * 1e3;
    printf("{\"bench\": \"device\", \"name\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"model_peak_timad_s\": %.3f}\n", p.name, sms,
           clock_khz / 1e3, model_peak / 1e12);
    for (int threads : {128, 256, 512, 1024}) {
        for (int per_sm : {1, 2, 4, 8}) {
            if (threads * per_sm > 2048) continue;
            const int grid = sms * per_sm;
            const int iters = 4096;
            float ms = time_kernel([&]() { k_imad_wide<<<grid, threads>>>(out, iters, 12345u); }, 5);
            double ops = (double)grid * threads * iters * 8.0;
            printf("{\"bench\": \"imad_wide\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"timad_s\": %.3f, \"frac_model\": %.3f}\n",
                   threads, per_sm, ms, ops / ms / 1e9, ops / (ms * 1e-3) / model_peak);
            ms = time_kernel([&]() { k_imad_chain<<<grid, threads>>>(out, iters, 12345u); }, 5);
            ops = (double)grid * threads * iters * 8.0;     // 2 chains x 4 fused lo/hi pairs
            printf("{\"bench\": \"imad_chain\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"timad_s\": %.3f, \"frac_model\": %.3f}\n",
                   threads, per_sm, ms, ops / ms / 1e9, ops / (ms * 1e-3) / model_peak);
        }
    }
#define FRB(ILP, SQR, NAME)                                                                                            \
    for (int threads : {128, 256}) {                                                                                   \
        for (int per_sm : {1, 2, 3, 4, 6, 8}) {                                                                        \
            if (threads * per_sm > 2048) continue;                                                                     \
            const int grid = sms * per_sm;                                                                             \
            const int iters = 2048;                                                                                    \
            float ms = time_kernel([&]() { k_fr_mul<ILP, SQR><<<grid, threads>>>(out, iters, 777u); }, 3);             \
            double fm = (double)grid * threads * iters * ILP;                                                          \
            printf("{\"bench\": \"%s\", \"ilp\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f, " \
                   "\"timad_s\": %.3f, \"frac_model\": %.3f}\n",                                                       \
                   NAME, ILP, threads, per_sm, ms, fm / ms / 1e6, fm * 136.0 / ms / 1e9, fm * 136.0 / (ms * 1e-3) / model_peak); \
        }                                                                                                              \
    }
#define FRB29(ILP, SQR, NAME)                                                                                          \
    for (int threads : {128, 256}) {                                                                                   \
        for (int per_sm : {1, 2, 4, 8}) {                                                                              \
            if (threads * per_sm > 2048) continue;                                                                     \
            const int grid = sms * per_sm;                                                                             \
            const int iters = 2048;                                                                                    \
            float ms = time_kernel([&]() { k_fr29<ILP, SQR><<<grid, threads>>>(out, iters, 777u); }, 3);               \
            double fm = (double)grid * threads * iters * ILP;                                                          \
            printf("{\"bench\": \"%s\", \"ilp\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n", \
                   NAME, ILP, threads, per_sm, ms, fm / ms / 1e6);                                                     \
        }                                                                                                              \
    }
    // sustained vs burst: the same fr_mul kernel for ~5 ms and for ~0.5 s (do clocks / power limit the long run?)
    for (int mult : {1, 8, 64, 128}) {
        const int threads = 256, grid = sms * 4, iters = 2048 * mult;
        float ms = time_kernel([&]() { k_fr_mul<1, false><<<grid, threads>>>(out, iters, 777u); }, 2);
        double fm = (double)grid * threads * iters;
        printf("{\"bench\": \"fr_mul_sustained\", \"ms\": %.2f, \"gfmul_s\": %.2f}\n", ms, fm / ms / 1e6);
    }
    // FP64-pipe multiplier alone, and CONCURRENTLY with the IMAD multiplier (two streams): do the pipes overlap?
    for (int threads : {128, 256}) {
        for (int per_sm : {2, 4, 8}) {
            if (threads * per_sm > 2048) continue;
            const int grid = sms * per_sm;
            const int iters = 2048;
            float ms = time_kernel([&]() { k_fr52<1><<<grid, threads>>>(out, iters, 777u); }, 3);
            double fm = (double)grid * threads * iters;
            printf("{\"bench\": \"fr52_dfma_mul\", \"ilp\": 1, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n",
                   threads, per_sm, ms, fm / ms / 1e6);
            ms = time_kernel([&]() { k_fr52<2><<<grid, threads>>>(out, iters, 777u); }, 3);
            printf("{\"bench\": \"fr52_dfma_mul\", \"ilp\": 2, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gfmul_s\": %.2f}\n",
                   threads, per_sm, ms, 2 * fm / ms / 1e6);
        }
    }
    {
        cudaStream_t s1, s2;
        CK(cudaStreamCreate(&s1));
        CK(cudaStreamCreate(&s2));
        uint32_t* out2;
        CK(cudaMalloc(&out2, 64));
        const int threads = 256, iters = 4096;
        for (int per_sm : {1, 2, 3}) {
            const int grid = sms * per_sm;
            float t_imad = time_kernel([&]() { k_fr_mul<1, false><<<grid, threads, 0, s1>>>(out, iters, 777u); }, 3);
            float t_dfma = time_kernel([&]() { k_fr52<1><<<grid, threads, 0, s2>>>(out2, iters, 777u); }, 3);
            // both at once: events on the legacy stream bracket both streams via device sync
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaDeviceSynchronize());
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0));
                CK(cudaEventCreate(&e1));
                CK(cudaEventRecord(e0, s1));
                CK(cudaStreamWaitEvent(s2, e0, 0));
                k_fr_mul<1, false><<<grid, threads, 0, s1>>>(out, iters, 777u);
                k_fr52<1><<<grid, threads, 0, s2>>>(out2, iters, 777u);
                cudaEvent_t j;
                CK(cudaEventCreate(&j));
                CK(cudaEventRecord(j, s2));
                CK(cudaStreamWaitEvent(s1, j, 0));
                CK(cudaEventRecord(e1, s1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            double fm = (double)grid * threads * iters;
            printf("{\"bench\": \"dual_pipe\", \"ctas_per_sm_each\": %d, \"ms_imad_alone\": %.3f, \"ms_dfma_alone\": %.3f, \"ms_both\": %.3f, "
                   "\"gfmul_s_imad_alone\": %.2f, \"gfmul_s_dfma_alone\": %.2f, \"gfmul_s_both\": %.2f}\n",
                   per_sm, t_imad, t_dfma, best, fm / t_imad / 1e6, fm / t_dfma / 1e6, 2 * fm / best / 1e6);
        }
    }
    FRB29(1, false, "fr29_mul")
    FRB29(2, false, "fr29_mul")
    FRB29(1, true, "fr29_sqr")
    FRB29(2, true, "fr29_sqr")
    FRB(1, false, "fr_mul")
    FRB(2, false, "fr_mul")
    FRB(4, false, "fr_mul")
    FRB(1, true, "fr_sqr")
    FRB(2, true, "fr_sqr")
    CK(cudaFree(out));
    return 0;
}
