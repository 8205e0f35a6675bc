// Microbenchmark: dependent-issue latency of scalar and packed FP32 ops, MUFU.LG2, LDS on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int MODE>
__global__ void kern(float *out, long long *cyc, float s) {
    __shared__ u64 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = pk((float)((i * 8 + 8) & 8191), 0.f);
    float a = s; u64 p = pk(s, s + 1.f), q = pk(s * 0.5f, s * 0.25f); unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 8;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) {
        if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(s));
        if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(s));
        if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
        if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));
        if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
        if (MODE == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a));
        if (MODE == 6) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr)); addr = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(v & 0) + threadIdx.x * 8 + (i & 1) * 8; p = v; }
        if (MODE == 7) { float x, y; up(p, x, y); u64 sw = pk(y, -x); asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(sw), "l"(q)); }  // swizzled dependent
        if (MODE == 8) { float x, y; up(p, x, y); asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a) : "f"(x), "f"(a)); p = pk(a, y); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q)); } // packed -> scalar -> packed
    }
    long long t1 = clock64();
    float x, y; up(p, x, y);
    out[threadIdx.x] = a + x + y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *name, int per) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    kern<MODE><<<1, 32>>>(out, cyc, 1.0001f); kern<MODE><<<1, 32>>>(out, cyc, 1.0001f);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %.2f cycles per op\n", name, (double)h / 256 / per);
}
int main() {
    run<0>("FADD dependent", 1); run<1>("FFMA dependent", 1); run<2>("FADD2 dependent", 1); run<3>("FFMA2 dependent", 1);
    run<4>("FMUL2 dependent", 1); run<5>("MUFU.LG2 dependent", 1); run<6>("LDS.64 dependent (pointer chase)", 1);
    run<7>("FADD2 swizzled dependent", 1); run<8>("FADD -> FADD2 chain (per pair)", 1);
    return 0;
}
