// Dependent-chain latency microbenchmark (sm_100a): cycles per link of a serial chain, one warp per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNR 16

template <int MODE>
__global__ void k(double *dout, long long *clk, double seed, int zero) {
    double d = seed + threadIdx.x, e = seed * 0.5, best = -1e300;
    float f = (float)seed + threadIdx.x;
    long long key = (long long)threadIdx.x + zero;
    const double g = seed * 0.25;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (MODE == 0) {                 // DADD chain
                d = d + g;
            } else if (MODE == 1) {          // DADD + DSETP + select (Viterbi relaxation chain through the max)
                const double c = d + g;
                d = c > e ? c : e;           // e is loop invariant: chain = DADD -> DSETP -> FSEL
                e = e + 0.0 * zero;
            } else if (MODE == 2) {          // DSETP + select only
                d = d > e ? d : e + (double)i;
            } else if (MODE == 3) {          // SHFL.UP of a double + DADD
                d = __shfl_up_sync(0xffffffffu, d, 1) + g;
            } else if (MODE == 4) {          // SHFL.UP of a double alone (2 SHFL)
                d = __shfl_up_sync(0xffffffffu, d, 1);
            } else if (MODE == 5) {          // FADD chain
                f = f + (float)g;
            } else if (MODE == 6) {          // scan round: SHFL + DADD + DSETP + select
                const double al = __shfl_up_sync(0xffffffffu, d, 1) + g;
                d = d >= al ? d : al;
            } else if (MODE == 7) {          // 64-bit integer compare + select chain (ordered keys)
                const long long c = key + 12345;
                key = c > (long long)i * 7 + zero ? c : (long long)i * 7 + zero;
            } else if (MODE == 8) {          // DMUL chain
                d = d * 1.0000001;
            } else if (MODE == 9) {          // LDS.128-dependent chain is not measured here
                d = fma(d, 1.0000001, g);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    dout[blockIdx.x * blockDim.x + threadIdx.x] = d + e + best + f + (double)key;
}

template <int MODE>
void run(const char *name) {
    double *dout; long long *clk;
    cudaMalloc(&dout, 148 * 32 * 8); cudaMalloc(&clk, 148 * 8);
    k<MODE><<<1, 32>>>(dout, clk, 1.5, 0);
    cudaDeviceSynchronize();
    k<MODE><<<1, 32>>>(dout, clk, 1.5, 0);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.2f cycles per link\n", name, (double)c / (ITERS * UNR));
    cudaFree(dout); cudaFree(clk);
}

int main() {
    run<0>("DADD");
    run<1>("DADD + DSETP + select");
    run<2>("DSETP + select (+DADD off chain)");
    run<3>("SHFL.UP(double) + DADD");
    run<4>("SHFL.UP(double)");
    run<5>("FADD");
    run<6>("scan round: SHFL + DADD + DSETP(>=) + select");
    run<7>("int64 add + compare + select");
    run<8>("DMUL");
    run<9>("DFMA");
    return 0;
}
