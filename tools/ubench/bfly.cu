// Microbenchmark: register-only radix-16 butterfly (+15 twiddle products) throughput per SM
// sub-partition, packed (default) vs scalar (-DRFA_NO_PACKED) FP32.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rfanalyzer_b200/csrc/rfa_fft_core.cuh"
using namespace rfa;
template <bool TW>
__global__ void __launch_bounds__(256, 2) kern(float *out, long long *cyc, int iters, float s) {
    cf u[16], tw[15];
    for (int i = 0; i < 16; i++) u[i] = cf{threadIdx.x * 0.001f + i, s * i};
    for (int i = 0; i < 15; i++) tw[i] = cf{cosf(0.01f * i * (threadIdx.x + 1)), sinf(0.01f * i * (threadIdx.x + 1))};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (TW) {
#pragma unroll
            for (int r = 1; r < 16; r++) u[r] = cmul(u[r], tw[r - 1]);
        }
        Dft<16>::run(u);
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 16; i++) acc += u[i].x + u[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <bool TW> void run(const char *name, int ctas_per_sm) {
    int sms = 148, iters = 2000; float *out; long long *cyc;
    cudaMalloc(&out, 4 * sms * ctas_per_sm * 256); cudaMalloc(&cyc, 8 * sms * ctas_per_sm);
    kern<TW><<<sms * ctas_per_sm, 256>>>(out, cyc, 10, 1e-3f);
    kern<TW><<<sms * ctas_per_sm, 256>>>(out, cyc, iters, 1e-3f);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, cyc, 8 * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms * ctas_per_sm; i++) avg += h[i]; avg /= sms * ctas_per_sm;
    int wps = 2 * ctas_per_sm;  // warps per SMSP
    printf("%-28s warps/SMSP=%d  cycles per butterfly per warp = %.1f   per SMSP = %.1f\n", name, wps, avg / iters, avg / iters / wps);
}
int main() {
#ifdef RFA_PACKED_OFF
    printf("scalar build\n");
#endif
    run<false>("radix-16 only", 1); run<false>("radix-16 only", 2);
    run<true>("15 cmul + radix-16", 1); run<true>("15 cmul + radix-16", 2);
    return 0;
}
