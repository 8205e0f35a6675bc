// Microbenchmark: issue rate of scalar vs packed (f32x2) FP32 instructions on sm_100a.
// Prints warp-instructions per clock per SM sub-partition for each variant.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE, int NACC>
__global__ void __launch_bounds__(1024) kern(float *out, long long *cyc, int iters, float s) {
    float a[NACC]; u64 p[NACC];
    for (int i = 0; i < NACC; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    float b = s, c = s * 0.5f; u64 pb = pk(b, b * 1.01f), pc = pk(c, c * 1.01f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                // two distinct accumulators as sources (butterfly-like a+b / a-b)
                if (MODE == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % NACC]));
                if (MODE == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) % NACC]));
                // fma with three distinct per-thread registers
                if (MODE == 8) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(a[(i + 1) % NACC]), "f"(a[(i + 2) % NACC]));
                if (MODE == 9) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i + 1) % NACC]), "l"(p[(i + 2) % NACC]));
            }
        }
    }
    long long t1 = clock64();
    float acc = 0; for (int i = 0; i < NACC; i++) acc += a[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int NACC>
void run(const char *name, int threads, int ctas_per_sm) {
    int sms = 148, iters = 2000; float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * threads); cudaMalloc(&cyc, 8 * sms * ctas_per_sm);
    kern<MODE, NACC><<<sms * ctas_per_sm, threads>>>(out, cyc, 10, 1.0001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<MODE, NACC><<<sms * ctas_per_sm, threads>>>(out, cyc, iters, 1.0001f);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8]; cudaMemcpy(h, cyc, 8 * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms * ctas_per_sm; i++) avg += h[i]; avg /= sms * ctas_per_sm;
    double winstr = (double)iters * 8 * NACC * (threads / 32) * ctas_per_sm;  // warp-instr per SM
    printf("%-34s thr=%4d cta/sm=%d  cycles=%9.0f  warp-instr/clk/SMSP=%.3f  (ms=%.3f, %.1f Gwinstr/s chip)\n", name, threads,
           ctas_per_sm, avg, winstr / avg / 4.0, ms, winstr * sms / ms / 1e6);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int thr : {256, 512, 1024}) {
        run<0, 8>("FFMA  (acc,b,c) 8acc", thr, 1);
        run<1, 8>("FFMA2 (acc,b,c) 8acc", thr, 1);
        run<2, 8>("FADD  (acc,b) 8acc", thr, 1);
        run<3, 8>("FADD2 (acc,b) 8acc", thr, 1);
        run<4, 8>("FMUL  (acc,b) 8acc", thr, 1);
        run<5, 8>("FMUL2 (acc,b) 8acc", thr, 1);
        run<6, 8>("FADD  (acc,acc') 8acc", thr, 1);
        run<7, 8>("FADD2 (acc,acc') 8acc", thr, 1);
        run<8, 8>("FFMA  (a',a'',acc) 8acc", thr, 1);
        run<9, 8>("FFMA2 (a',a'',acc) 8acc", thr, 1);
    }
    run<0, 16>("FFMA  16acc", 512, 1);
    run<1, 16>("FFMA2 16acc", 512, 1);
    run<7, 16>("FADD2 (acc,acc') 16acc", 512, 1);
    return 0;
}
