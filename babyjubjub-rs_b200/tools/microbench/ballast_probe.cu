// Compile-only experiment behind csrc/fr.cuh::fma_ballast: how ptxas distributes integer adds / moves between the ALU
// and the fma pipe INSIDE an out-of-line multiplication subroutine depends on the instruction mix of the whole kernel.
//   VARIANT 0: trivial caller            VARIANT 1: ALU-heavy caller (400 shift/xor/add steps per iteration)
//   VARIANT 2: the same ALU-heavy caller + 1,500 multiply-adds behind a branch that is never taken (the ballast)
// tools/microbench/ballast_probe.sh compiles the three to cubins and counts, per subroutine, IMAD.WIDE and the other
// fma-pipe instructions ("passengers": IMAD.MOV / IMAD.IADD / IMAD.X / IMAD.SHL / HFMA2).  Nothing here runs.
#include "vm.cuh"
using namespace bjj;
using namespace bjj::vm;
#ifndef VARIANT
#define VARIANT 0
#endif
extern "C" __global__ void __launch_bounds__(128, 5) probe(int n, int* out, const uint32_t* in) {
    Slot a = slot(0), b = slot(1), c = slot(2), d = slot(3), e = slot(4), f = slot(5);
    uint32_t h = in[threadIdx.x];
    for (int i = 0; i < n; i++) {
        mul2(a, b, c, d, e, f);
        mul2(b, a, d, e, c, f);
        mul(a, b, c);
#if VARIANT >= 1
#pragma unroll
        for (int j = 0; j < 400; j++) {
            h = (h << 5) ^ (h >> 3) ^ (h + j * 77u);
            h += __popc(h);
        }
#endif
    }
#if VARIANT >= 2
    if (n == -12345) {   // never
        uint32_t g = h;
#pragma unroll
        for (int j = 0; j < 1500; j++) g = g * (h | 3) + in[j & 7];
        h = g;
    }
#endif
    out[threadIdx.x] = ld_word(a, 0) + h;
}
