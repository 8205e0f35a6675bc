// Microbenchmark: can other pipes issue in the shadow of packed FP32 (FADD2/FFMA2) on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

// MODE: 0 = 8 FADD2 only; 1 = 8 FADD2 + 8 LOP3; 2 = 8 FADD2 + 8 LDS.64; 3 = 8 FADD2 + 8 FADD (scalar)
//       4 = 8 FADD2 + 8 LOP3 + 4 LDS.64; 5 = 8 swizzled FADD2 (LO_HI.NP); 6 = 8 FFMA2 with scalar broadcast
//       7 = 8 FADD2 + 2 MUFU.LG2; 8 = 16 FADD only; 9 = 16 FADD + 8 LOP3;  10 = 8 FADD2 + 8 STS.64
template <int MODE>
__global__ void __launch_bounds__(512) kern(float *out, long long *cyc, int iters, float s) {
    __shared__ u64 sm[2048];
    u64 p[8]; unsigned q[8]; float a[16]; u64 l[8];
    for (int i = 0; i < 8; i++) { p[i] = pk(threadIdx.x * 0.001f + i, 0.5f + i); q[i] = threadIdx.x + i; l[i] = 0; }
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.01f + i;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = pk(i, i);
    u64 pb = pk(s, s * 1.01f); float w = s * 0.999f; float lg[2] = {s + 3.f, s + 4.f};
    unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 8;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE <= 4 || MODE == 7 || MODE == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 1 || MODE == 4 || MODE == 9) asm volatile("lop3.b32 %0, %0, %1, 0x55555555, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
                if (MODE == 2 || (MODE == 4 && (i & 1))) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(l[i]) : "r"(sa + (i * 4096 % 8192 + (i / 2) * 8)));
                if (MODE == 10) asm volatile("st.shared.b64 [%1], %0;" :: "l"(p[(i + 4) & 7]), "r"(sa + (i * 4096 % 8192)));
                if (MODE == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(w));
                if (MODE == 5) {
                    float x, y; up(p[(i + 1) & 7], x, y);
                    u64 sw = pk(y, -x);
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(sw));
                }
                if (MODE == 6) {
                    float x, y; up(p[(i + 1) & 7], x, y);
                    u64 sw = pk(-y, x), ww = pk(w, w);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(sw), "l"(ww));
                }
                if (MODE == 7 && (i & 3) == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(lg[i >> 2]));
                if (MODE == 8 || MODE == 9) {
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(w));
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i + 8]) : "f"(w));
                }
            }
        }
    }
    long long t1 = clock64();
    float acc = lg[0] + lg[1];
    for (int i = 0; i < 8; i++) { float x, y; up(p[i], x, y); acc += x + y + q[i]; up(l[i], x, y); acc += x + y; }
    for (int i = 0; i < 16; i++) acc += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads) {
    int sms = 148, iters = 2000; float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads); cudaMalloc(&cyc, 8 * sms);
    kern<MODE><<<sms, threads>>>(out, cyc, 10, 1.0001f);
    kern<MODE><<<sms, threads>>>(out, cyc, iters, 1.0001f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    // cycles per SMSP per (group of 8 slots) per warp
    double groups = (double)iters * 4 * (threads / 32) / 4.0;
    printf("%-44s thr=%4d  cycles per 8-slot group per warp (SMSP) = %.2f\n", name, threads, avg / groups);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int thr : {256, 512}) {
        run<0>("8 FADD2", thr);
        run<1>("8 FADD2 + 8 LOP3", thr);
        run<2>("8 FADD2 + 8 LDS.64", thr);
        run<10>("8 FADD2 + 8 STS.64", thr);
        run<3>("8 FADD2 + 8 FADD", thr);
        run<4>("8 FADD2 + 8 LOP3 + 4 LDS.64", thr);
        run<5>("8 FADD2 swizzled LO_HI.NP", thr);
        run<6>("8 FFMA2 swizzled + scalar bcast", thr);
        run<7>("8 FADD2 + 2 MUFU.LG2", thr);
        run<8>("16 FADD", thr);
        run<9>("16 FADD + 8 LOP3", thr);
    }
    return 0;
}
