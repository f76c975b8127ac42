// Column body of the linear-gap alignment scan, scalar against packed (add.rn.f32x2) adds: cycles per 30-row column
// per warp with 16 single-warp CTAs per SM (the scan kernel's launch shape).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o cell cell.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R 30
#define S 6
#define K 5
#define COLS 4096

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(32) k(float *out, const float *lut, long long *clk, float gh, float gv) {
    float Sv[R];
    for (int r = 0; r < R; ++r) Sv[r] = -(float)(r + threadIdx.x);
    float botS = 0.f, diag_next = 0.f;
    const float *row = lut + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < COLS; ++j) {
        float sc[K + 1];
#pragma unroll
        for (int kk = 0; kk <= K; ++kk) sc[kk] = __ldg(row + ((j * 7 + kk) & 255) * 32);
        float inS = __shfl_up_sync(0xffffffffu, botS, 1);
        if (threadIdx.x == 0) inS = 0.f;
        float diag = diag_next;
        diag_next = inS;
        float cS = inS;
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float pS = Sv[r];
                const float inter = diag + sc[r / S];
                diag = pS;
                cS = fmax3(inter, pS + gh, cS + gv);
                Sv[r] = cS;
            }
        } else {
            // shifted mapping: local row 0 belongs to the previous lane's last level (sc[K]); rows 1..29: level (r-1)/6
            const unsigned long long gh2 = pack(gh, gh);
            unsigned long long hp[R / 2], ip[R / 2];
#pragma unroll
            for (int m = 0; m < R / 2; ++m) {
                const unsigned long long P = pack(Sv[2 * m], Sv[2 * m + 1]);
                hp[m] = add2(P, gh2);
                ip[m] = add2(P, pack(sc[m / 3], sc[m / 3]));
            }
            float inter = diag + sc[K];
#pragma unroll
            for (int m = 0; m < R / 2; ++m) {
                float h0, h1, i1, i2;
                unpack(hp[m], h0, h1);
                unpack(ip[m], i1, i2);
                cS = fmax3(inter, h0, cS + gv);
                Sv[2 * m] = cS;
                cS = fmax3(i1, h1, cS + gv);
                Sv[2 * m + 1] = cS;
                inter = i2;
            }
        }
        botS = cS;
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += Sv[r];
    out[blockIdx.x * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ctas_per_sm) {
    const int blocks = 148 * ctas_per_sm;
    float *out, *lut; long long *clk;
    cudaMalloc(&out, blocks * 32 * 4); cudaMalloc(&lut, 256 * 32 * 4); cudaMalloc(&clk, blocks * 8);
    cudaMemset(lut, 0, 256 * 32 * 4);
    k<MODE><<<blocks, 32>>>(out, lut, clk, -1.f, -16.f);
    k<MODE><<<blocks, 32>>>(out, lut, clk, -1.f, -16.f);
    cudaDeviceSynchronize();
    long long *h = new long long[blocks];
    cudaMemcpy(h, clk, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // per SMSP: ctas_per_sm / 4 warps share a scheduler
    printf("%-28s %2d warps/SM: %.1f cycles per column per warp, %.1f per column per SMSP, %.2f clk per cell-row\n", name, ctas_per_sm,
           avg / COLS, avg / COLS / (ctas_per_sm / 4.0), avg / COLS / (ctas_per_sm / 4.0) / R);
    cudaFree(out); cudaFree(lut); cudaFree(clk); delete[] h;
}

int main() {
    for (int c : {16, 24, 28, 32}) {
        run<0>("scalar 3 FADD + FMNMX3", c);
        run<1>("packed 2 FADD2/2 + FADD", c);
    }
    return 0;
}
