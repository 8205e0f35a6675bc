#include "../../rfanalyzer_b200/csrc/fir.cu"  // build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rfanalyzer_b200/csrc -o build/dbg/stage_dbg tools/dbg/stage_dbg.cu
#include <cstdio>
#include <vector>
#include <cmath>
using namespace rfa;
template <int KIND>
__global__ void k_fast(StreamSrc src, long long k_al, int span, float2 *out) {
    extern __shared__ float2 sm[];
    float2 *xs = sm, *s_nco = sm + span + 8;
    fill_nco_pairs<KIND>(src, s_nco);
    __syncthreads();
    const int t0 = (int)(((long long)src.nco_idx + k_al) % src.nco_len);
    stage_span_pairs<KIND>(src, k_al, span, xs, s_nco, t0);
    __syncthreads();
    for (int i = threadIdx.x; i < (span & ~1); i += blockDim.x) out[i] = xs[i];
}
template <int KIND>
__global__ void k_slow(StreamSrc src, long long k_al, int span, float2 *out) {
    extern __shared__ float2 sm[];
    stage_span<KIND>(src, k_al, span, sm);
    __syncthreads();
    for (int i = threadIdx.x; i < (span & ~1); i += blockDim.x) out[i] = sm[i];
}
template <int KIND>
void run(int bps) {
    const int span = 3000, nco_len = 10, total = 8192;
    std::vector<unsigned char> raw(total * bps);
    for (size_t i = 0; i < raw.size(); i++) raw[i] = (unsigned char)((i * 2654435761u) >> 13);
    std::vector<float> c(nco_len), s(nco_len);
    for (int i = 0; i < nco_len; i++) { c[i] = (float)cos(2 * M_PI * i / nco_len); s[i] = (float)sin(2 * M_PI * i / nco_len); }
    unsigned char *d_raw; float *d_c, *d_s; float2 *o1, *o2;
    cudaMalloc(&d_raw, raw.size()); cudaMalloc(&d_c, 4 * nco_len); cudaMalloc(&d_s, 4 * nco_len);
    cudaMalloc(&o1, 8 * span); cudaMalloc(&o2, 8 * span);
    cudaMemcpy(d_raw, raw.data(), raw.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_c, c.data(), 4 * nco_len, cudaMemcpyHostToDevice); cudaMemcpy(d_s, s.data(), 4 * nco_len, cudaMemcpyHostToDevice);
    StreamSrc src{}; src.raw = d_raw; src.nco_cos = d_c; src.nco_sin = d_s; src.nco_len = nco_len; src.nco_idx = 3;
    const long long k_al = 16;
    k_fast<KIND><<<1, 256, (span + 8 + 64) * 8>>>(src, k_al, span, o1);
    k_slow<KIND><<<1, 256, (span + 8) * 8>>>(src, k_al, span, o2);
    std::vector<float2> h1(span), h2(span);
    cudaMemcpy(h1.data(), o1, 8 * span, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), o2, 8 * span, cudaMemcpyDeviceToHost);
    printf("KIND %d: %s\n", KIND, cudaGetErrorString(cudaDeviceSynchronize()));
    int bad = 0;
    for (int i = 0; i < (span & ~1); i++)
        if (memcmp(&h1[i], &h2[i], 8)) {
            if (bad < 8) {
                const int t = (int)((3 + k_al + i) % nco_len);
                printf("  i=%d t=%d fast=(%.9g,%.9g) slow=(%.9g,%.9g) c=%.9g s=%.9g raw=%02x %02x %02x %02x\n", i, t, h1[i].x, h1[i].y, h2[i].x, h2[i].y, c[t], s[t],
                       raw[(k_al + i) * bps], raw[(k_al + i) * bps + 1], bps == 4 ? raw[(k_al + i) * bps + 2] : 0, bps == 4 ? raw[(k_al + i) * bps + 3] : 0);
            }
            bad++;
        }
    printf("  mismatches: %d of %d\n", bad, span & ~1);
}
int main() { run<0>(2); run<1>(2); run<2>(4); return 0; }
