// Micro-benchmark: register-file read bandwidth.  ALU (VIADDMNMX) + FMA (IMAD) pairs whose source operands are
// 1, 2 or 3 DISTINCT registers each (rotating over a pool of 16, so nothing can sit in the operand reuse cache).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP 32
template <int MODE>
__global__ void k(uint32_t *out, long long *cyc, uint32_t seed)
{
    uint32_t x[16], y[16];
    for (int i = 0; i < 16; ++i) { x[i] = seed + i * 77 + threadIdx.x; y[i] = seed * 5 + i; }
    uint32_t b = seed * 3 + 1, c = seed ^ 0x55aa;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int i1 = (i + 5) & 15, i2 = (i + 11) & 15;
                if (MODE == 0) { x[i] = __viaddmin_u16x2(x[i], b, c); y[i] = y[i] * b + c; }                 // 1 + 1 fresh reads
                if (MODE == 1) { x[i] = __viaddmin_u16x2(x[i], x[i1], c); y[i] = y[i] * b + y[i1]; }          // 2 + 2
                if (MODE == 2) { x[i] = __viaddmin_u16x2(x[i], x[i1], x[i2]); y[i] = y[i] * y[i1] + y[i2]; }  // 3 + 3
                if (MODE == 3) { x[i] = __viaddmin_u16x2(x[i], x[i1], x[i2]); }                              // ALU only, 3
                if (MODE == 4) { y[i] = y[i] * y[i1] + y[i2]; }                                               // FMA only, 3
                if (MODE == 5) { x[i] = __viaddmin_u16x2(x[i], x[i1], x[i2]); y[i] = y[i] + y[i1]; }          // 3 + 2 (the ACS mix)
                if (MODE == 6) { x[i] = __viaddmin_u16x2(x[i], x[i1], c); y[i] = y[i] + y[i1]; }              // 2 + 2
                if (MODE == 7) { x[i] = __viaddmin_u16x2(x[i], b, c); y[i] = y[i] + b; }                      // 1 + 1, add
            }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 16; ++i) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char *name, uint32_t *out, long long *cyc)
{
    for (int warps : {1, 8}) {
        long long h;
        k<MODE><<<1, 32 * warps>>>(out, cyc, 12345); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SM %d: %.2f cycles per (pair | instr) per warp\n", name, warps, (double)h / (64.0 * REP * 16));
    }
}
int main()
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    run<0>("addmin(1 fresh) + imad(1 fresh)", out, cyc);
    run<1>("addmin(2 fresh) + imad(2 fresh)", out, cyc);
    run<2>("addmin(3 fresh) + imad(3 fresh)", out, cyc);
    run<3>("addmin(3 fresh) alone", out, cyc);
    run<4>("imad(3 fresh) alone", out, cyc);
    run<5>("addmin(3 fresh) + add(2 fresh)", out, cyc);
    run<6>("addmin(2 fresh) + add(2 fresh)", out, cyc);
    run<7>("addmin(1 fresh) + add(1 fresh)", out, cyc);
    return 0;
}
