// synth.cu -- deterministic synthetic IQ generator on the device (benchmark/test input).
//
// The reference ships no input fixtures beyond inline sinusoids (ApplicationTest.kt:33-38);
// benchmarks use the all-integer generator of SURVEY.md 8(d): hashed noise plus phase-
// accumulator tones read from a 4096-entry cosine table, so a recording of any length can be
// produced in place on each GPU (sample n depends on n alone) and reproduced bit for bit on
// the CPU.  One thread per sample pair, 32-bit stores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

namespace rfa {
namespace {

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

struct SynthArgs {
    SynthComp comp[8];
    int ncomp;
    int noise_shift;
    uint32_t seed;
    unsigned long long first;
    long long nsamples;
};

template <int FMT>
__device__ __forceinline__ void one_sample(const SynthArgs &a, const int16_t *tab, unsigned long long n, int &vi,
                                           int &vq) {
    const uint32_t h = fmix32(a.seed ^ (uint32_t)n ^ ((uint32_t)(n >> 32) * 0x9E3779B9u));
    if (FMT == 2) {
        vi = ((int)(int16_t)(h & 0xFFFF)) >> a.noise_shift;
        vq = ((int)(int16_t)(h >> 16)) >> a.noise_shift;
    } else {
        vi = ((int)(int8_t)(h & 0xFF)) >> a.noise_shift;
        vq = ((int)(int8_t)((h >> 8) & 0xFF)) >> a.noise_shift;
    }
    for (int c = 0; c < a.ncomp; c++) {
        uint32_t ph = (uint32_t)(n * a.comp[c].step);
        if (a.comp[c].mod_k != 0) {
            const uint32_t mph = (uint32_t)(n * a.comp[c].mod_step);
            const int m = tab[((mph - 0x40000000u) >> 20) & 4095];
            ph += (uint32_t)((long long)a.comp[c].mod_k * (long long)m);
        }
        const int ci = tab[(ph >> 20) & 4095];
        const int si = tab[((ph - 0x40000000u) >> 20) & 4095];
        vi += (a.comp[c].amp * ci + 8192) >> 14;
        vq += (a.comp[c].amp * si + 8192) >> 14;
    }
    const int lo = FMT == 2 ? -32768 : -128, hi = FMT == 2 ? 32767 : 127;
    vi = min(max(vi, lo), hi);
    vq = min(max(vq, lo), hi);
}

template <int FMT>
__global__ void __launch_bounds__(256) synth_kernel(const SynthArgs a, const int16_t *__restrict__ tab_g,
                                                    void *__restrict__ out) {
    __shared__ int16_t tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = tab_g[i];
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (FMT == 2) {
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.nsamples; k += stride) {
            int vi, vq;
            one_sample<FMT>(a, tab, a.first + (unsigned long long)k, vi, vq);
            ((uint32_t *)out)[k] = (uint32_t)(uint16_t)(int16_t)vi | ((uint32_t)(uint16_t)(int16_t)vq << 16);
        }
    } else {
        const int bias = FMT == 1 ? 128 : 0;
        // two samples per thread: one 32-bit store
        const long long pairs = a.nsamples / 2;
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < pairs; k += stride) {
            int i0, q0, i1, q1;
            one_sample<FMT>(a, tab, a.first + (unsigned long long)(2 * k), i0, q0);
            one_sample<FMT>(a, tab, a.first + (unsigned long long)(2 * k + 1), i1, q1);
            ((uint32_t *)out)[k] = (uint32_t)((i0 + bias) & 0xFF) | ((uint32_t)((q0 + bias) & 0xFF) << 8) |
                                   ((uint32_t)((i1 + bias) & 0xFF) << 16) | ((uint32_t)((q1 + bias) & 0xFF) << 24);
        }
        if ((a.nsamples & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            int i0, q0;
            one_sample<FMT>(a, tab, a.first + (unsigned long long)(a.nsamples - 1), i0, q0);
            ((uint16_t *)out)[a.nsamples - 1] = (uint16_t)(((i0 + bias) & 0xFF) | (((q0 + bias) & 0xFF) << 8));
        }
    }
}

}  // namespace

void synth_make_table(short *tab) {
    const double PI = 3.14159265358979323846;
    for (int j = 0; j < 4096; j++) tab[j] = (short)lround(16384.0 * cos(2.0 * PI * j / 4096.0));
}

cudaError_t synth_launch(int fmt, unsigned int seed, const SynthComp *comps, int ncomp, int noise_shift,
                         unsigned long long first, long long nsamples, const short *tab_dev, void *out,
                         int num_sms, cudaStream_t st) {
    if (nsamples <= 0) return cudaSuccess;
    if (ncomp < 0 || ncomp > 8) return cudaErrorInvalidValue;
    SynthArgs a{};
    for (int i = 0; i < ncomp; i++) a.comp[i] = comps[i];
    a.ncomp = ncomp;
    a.noise_shift = noise_shift;
    a.seed = seed;
    a.first = first;
    a.nsamples = nsamples;
    long long work = fmt == 2 ? nsamples : (nsamples + 1) / 2;
    long long blocks = (work + 255) / 256;
    if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
    if (blocks < 1) blocks = 1;
    switch (fmt) {
        case 0: synth_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(a, tab_dev, out); break;
        case 1: synth_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a, tab_dev, out); break;
        case 2: synth_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a, tab_dev, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace rfa
