// Pipe-rate microbenchmark, integer / DPX side (sm_100a): warp-instructions per clock per SMSP for the
// instruction mixes a fixed-point (tagged int32) Viterbi column is built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes_int pipes_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 16

template <int MODE>
__global__ void k(int *out, float *fout, long long *clk, int seed, float fseed) {
    int a[UNR];
    float f[UNR];
    for (int i = 0; i < UNR; ++i) { a[i] = seed * (i + 3) + threadIdx.x; f[i] = fseed + i + threadIdx.x; }
    const int g = seed * 5 + 1, gb = seed * 7 + 2, gc = seed * 11 + 3, gm = seed | 1;
    const float fg = fseed * 0.5f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            const int j = (i + 1) % UNR, l = (i + 2) % UNR;
            if (MODE == 0) {            // VIADDMNMX reg,reg,reg
                a[i] = max(a[j] + g, a[i]);
            } else if (MODE == 1) {     // VIMNMX3
                a[i] = max(max(a[i], a[j]), a[l] ^ it);
            } else if (MODE == 2) {     // IADD3 / VIADD
                a[i] = a[i] + a[j] + g;
            } else if (MODE == 3) {     // IMAD
                a[i] = a[j] * gm + a[i];
            } else if (MODE == 4) {     // LOP3
                a[i] = (a[i] & gm) ^ a[j];
            } else if (MODE == 5) {     // F2I
                a[i] = __float2int_rn(f[i]) ; f[i] = __int_as_float(a[i] ^ a[j]);
            } else if (MODE == 6) {     // M state: 1 add + 6 add-max + mask + add  (9 instr)
                int b = a[i] + g;
                b = max(a[j] + gb, b);
                b = max(a[l] + gc, b);
                b = max(a[(i + 3) % UNR] + g, b);
                b = max(a[(i + 4) % UNR] + gb, b);
                b = max(a[(i + 5) % UNR] + gc, b);
                b = max(a[(i + 6) % UNR] + g, b);
                a[i] = (b & ~7) + a[(i + 7) % UNR];
            } else if (MODE == 7) {     // M state as a tree: 7 adds + 3 max3 + mask + add (12 instr)
                int c0 = a[i] + g, c1 = a[j] + gb, c2 = a[l] + gc, c3 = a[(i + 3) % UNR] + g, c4 = a[(i + 4) % UNR] + gb,
                    c5 = a[(i + 5) % UNR] + gc, c6 = a[(i + 6) % UNR] + g;
                int b = max(max(c0, c1), c2);
                b = max(max(b, c3), c4);
                b = max(max(b, c5), c6);
                a[i] = (b & ~7) + a[(i + 7) % UNR];
            } else if (MODE == 8) {     // add-max next to float work on the fma pipe: 2 VIADDMNMX + 2 FFMA
                a[i] = max(a[j] + g, a[i]);
                f[i] = f[i] * fg + f[j];
                a[j] = max(a[l] + gb, a[j]);
                f[j] = f[j] * fg + f[l];
            } else if (MODE == 9) {     // FSETP + SEL (compare bit) + FMNMX
                const bool gt = f[j] > f[i];
                f[i] = fmaxf(f[i], f[j]);
                a[i] = gt ? (a[i] | 4) : a[i];
            } else if (MODE == 10) {    // PRMT
                a[i] = __byte_perm(a[i], a[j], 0x4240);
            } else if (MODE == 11) {    // SHFL (32-bit)
                a[i] = __shfl_up_sync(0xffffffffu, a[i], 1) + 1;
            } else if (MODE == 12) {    // FADD + FMNMX (float edge, no arg)
                f[i] = fmaxf(f[j] + fg, f[i]);
            } else if (MODE == 13) {    // IMAD-add (a * 1 + b kept as IMAD through a runtime 1) + VIMNMX3: adds on fma pipe
                int c0 = a[i] * gm + g, c1 = a[j] * gm + gb;
                a[i] = max(max(c0, c1), a[l]);
            }
        }
    }
    long long t1 = clock64();
    int s = 0; float fs = 0.f;
    for (int i = 0; i < UNR; ++i) { s ^= a[i]; fs += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    fout[blockIdx.x * blockDim.x + threadIdx.x] = fs;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_slot, int warps) {
    int *out; float *fout; long long *clk;
    const int blocks = 148;
    cudaMalloc(&out, blocks * warps * 32 * 4); cudaMalloc(&fout, blocks * warps * 32 * 4); cudaMalloc(&clk, blocks * 8);
    k<MODE><<<blocks, warps * 32>>>(out, fout, clk, 3, 1.25f);
    k<MODE><<<blocks, warps * 32>>>(out, fout, clk, 3, 1.25f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double slots = (double)ITERS * UNR * warps;      // per SM
    printf("%-52s warps/SM %2d: %.3f cycles per slot per SMSP (%d instr/slot -> %.3f warp-instr/clk/SMSP)\n", name, warps,
           avg / (slots / 4), instr_per_slot, instr_per_slot * (slots / 4) / avg);
    cudaFree(out); cudaFree(fout); cudaFree(clk);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("VIADDMNMX", 1, w);
        run<1>("VIMNMX3 (+LOP)", 2, w);
        run<2>("IADD3", 1, w);
        run<3>("IMAD", 1, w);
        run<4>("LOP3", 1, w);
        run<5>("F2I + LOP3", 2, w);
        run<6>("M state chain: IADD + 6 VIADDMNMX + LOP3 + IADD", 9, w);
        run<7>("M state tree: 7 IADD + 3 VIMNMX3 + LOP3 + IADD", 12, w);
        run<8>("2 VIADDMNMX + 2 FFMA", 4, w);
        run<9>("FSETP + SEL + FMNMX", 3, w);
        run<10>("PRMT", 1, w);
        run<11>("SHFL + IADD", 2, w);
        run<12>("FADD + FMNMX", 2, w);
        run<13>("2 IMAD + VIMNMX3", 3, w);
    }
    return 0;
}
