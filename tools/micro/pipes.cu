// Pipe-rate microbenchmark (sm_100a): warp-instructions per clock per SM for the instruction mixes the DP
// kernels are built from.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 16

template <int MODE>
__global__ void k(float *out, double *dout, long long *clk, float seed) {
    float a[UNR];
    float2 p[UNR];
    double d[UNR];
    for (int i = 0; i < UNR; ++i) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); d[i] = a[i]; }
    const float g = seed * 0.5f;
    const float2 g2 = make_float2(g, g + 1.f);
    const double gd = g;
    const float gb = seed * 0.25f, gc = seed * 0.125f;
    const float2 g2b = make_float2(gb, gb + 1.f), g2c = make_float2(gc, gc + 1.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (MODE == 0) {            // FADD
                a[i] = a[i] + g;
            } else if (MODE == 1) {     // FADD2
                unsigned long long r, x = *(unsigned long long *)&p[i], y = *(const unsigned long long *)&g2;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
                *(unsigned long long *)&p[i] = r;
            } else if (MODE == 2) {     // FMNMX3
                float r;
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a[i]), "f"(a[(i + 1) % UNR]), "f"(g));
                a[i] = r;
            } else if (MODE == 3) {     // FMNMX (2-input)
                a[i] = fmaxf(a[i], a[(i + 1) % UNR] );
            } else if (MODE == 4) {     // DADD
                d[i] = d[i] + gd;
            } else if (MODE == 5) {     // DSETP + select (fmax on double)
                d[i] = d[i] > d[(i + 1) % UNR] ? d[i] : d[(i + 1) % UNR];
            } else if (MODE == 6) {     // 3 FADD + 1 FMNMX3 (the linear cell)
                float x = a[i] + g, y = a[(i + 1) % UNR] + gb, z = a[(i + 2) % UNR] + gc;
                float r;
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z));
                a[i] = r;
            } else if (MODE == 7) {     // 3 FADD2 + 2 FMNMX3 (two packed cells)
                unsigned long long x, y, z, pi = *(unsigned long long *)&p[i], pj = *(unsigned long long *)&p[(i + 1) % UNR],
                                   pk = *(unsigned long long *)&p[(i + 2) % UNR], gg = *(const unsigned long long *)&g2, ggb = *(const unsigned long long *)&g2b, ggc = *(const unsigned long long *)&g2c;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(pi), "l"(gg));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(pj), "l"(ggb));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(pk), "l"(ggc));
                float2 fx = *(float2 *)&x, fy = *(float2 *)&y, fz = *(float2 *)&z, r;
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r.x) : "f"(fx.x), "f"(fy.x), "f"(fz.x));
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r.y) : "f"(fx.y), "f"(fy.y), "f"(fz.y));
                p[i] = r;
            } else if (MODE == 8) {     // DADD + compare + select (one Viterbi edge)
                double c = d[(i + 1) % UNR] + gd;
                d[i] = c > d[i] ? c : d[i];
            } else if (MODE == 9) {     // 64-bit integer key max (ordered keys): IADD.64 + ISETP.64 + 2 SEL
                long long x = __double_as_longlong(d[i]), y = __double_as_longlong(d[(i + 1) % UNR]);
                d[i] = __longlong_as_double(x > y ? x : y);
            } else if (MODE == 10) {    // DADD + DMNMX-like via fmax()
                d[i] = fmax(d[(i + 1) % UNR] + gd, d[i]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f; double sd = 0.0;
    for (int i = 0; i < UNR; ++i) { s += a[i] + p[i].x + p[i].y; sd += d[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    dout[blockIdx.x * blockDim.x + threadIdx.x] = sd;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_slot, int warps) {
    float *out; double *dout; long long *clk;
    const int blocks = 148;
    cudaMalloc(&out, blocks * warps * 32 * 4); cudaMalloc(&dout, blocks * warps * 32 * 8); cudaMalloc(&clk, blocks * 8);
    k<MODE><<<blocks, warps * 32>>>(out, dout, clk, 1.25f);
    k<MODE><<<blocks, warps * 32>>>(out, dout, clk, 1.25f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double slots = (double)ITERS * UNR * warps;      // per SM
    printf("%-44s warps/SM %2d: %.3f cycles per slot per SMSP (%d instr/slot -> %.3f warp-instr/clk/SMSP)\n", name, warps,
           avg / (slots / 4), instr_per_slot, instr_per_slot * (slots / 4) / avg);
    cudaFree(out); cudaFree(dout); cudaFree(clk);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FADD", 1, w);
        run<1>("FADD2 (add.f32x2)", 1, w);
        run<2>("FMNMX3", 1, w);
        run<3>("FMNMX", 1, w);
        run<4>("DADD", 1, w);
        run<5>("DSETP+select", 3, w);
        run<6>("3 FADD + FMNMX3", 4, w);
        run<7>("3 FADD2 + 2 FMNMX3 (2 cells)", 5, w);
        run<8>("DADD + DSETP + 2 SEL", 4, w);
        run<9>("int64 max", 4, w);
        run<10>("DADD + fmax", 4, w);
    }
    return 0;
}
