// fir.cu -- K4 rational polyphase resampler (optionally fused with K2, the converting NCO
// mixer) and K5 decimating FIR filters.
//
//   K5 = FirFilter.filter / filterReal          (A/dsp/FirFilter.kt:63-110, :121-163)
//        ComplexFirFilter.filter                (A/dsp/ComplexFirFilter.java:123-170)
//   K4 = RationalResampler.resample             (A/dsp/RationalResampler.kt:90-156)
//   K2 = IQConverter.mixPacketIntoSamplePacket  (A/source/*IQConverter*: see convert.cu)
//
// The reference walks a circular delay line sample by sample.  As a function of the sample
// STREAM that is: output j of a call is a dot product of the taps with the window of the
// stream that ends at input index i_j, with i_j given in closed form (FIR: first + j*dec;
// resampler: rel + floor((ph0 + j*D)/I), phase (ph0 + j*D) mod I).  State between calls is
// the last ntaps-1 stream samples ("history") plus two counters, kept by the host objects.
//
// Every CTA stages the input span of its output tile in shared memory once (decoding and
// mixing raw IQ bytes on the way when fused with K2), then each thread accumulates one output
// in the reference's tap order.  SUM_EXACT rounds every product and every sum separately
// (no FMA), which reproduces the JVM's float32 results bit for bit; SUM_FMA fuses them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "rfa_fft_core.cuh"

namespace rfa {
namespace {

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float a, float b) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(a, b)) : fmaf(a, b, acc);
}

// ---- input of the resampler: planar floats, or raw IQ codes mixed on the fly -------------
struct StreamSrc {
    const float *re, *im;      // planar floats (kind 3) -- or nullptr
    const void *raw;           // raw IQ codes (kind 0..2 = FMT_*)
    const float *hist_re, *hist_im;  // the `hist` samples that precede index 0
    int hist;
    const float *nco_cos, *nco_sin;  // device tables (nullptr = no mixing)
    int nco_len, nco_idx;            // table index of stream sample 0
};

template <int KIND>
__device__ __forceinline__ void fetch(const StreamSrc &s, long long k, float &re, float &im) {
    if (k < 0) {
        const long long h = (long long)s.hist + k;
        re = h >= 0 ? s.hist_re[h] : 0.0f;
        im = (h >= 0 && s.hist_im) ? s.hist_im[h] : 0.0f;
        return;
    }
    if (KIND == 3) {
        re = s.re[k];
        im = s.im ? s.im[k] : 0.0f;
        return;
    }
    float r, q;
    if (KIND == FMT_S8) {
        const uint16_t raw = ((const uint16_t *)s.raw)[k];
        r = conv_s8((int)(int8_t)(raw & 0xFF));
        q = conv_s8((int)(int8_t)(raw >> 8));
    } else if (KIND == FMT_U8) {
        const uint16_t raw = ((const uint16_t *)s.raw)[k];
        r = conv_u8((int)(raw & 0xFF));
        q = conv_u8((int)(raw >> 8));
    } else {
        const uint32_t raw = ((const uint32_t *)s.raw)[k];
        r = conv_s16((int)(int16_t)(raw & 0xFFFF));
        q = conv_s16((int)(int16_t)(raw >> 16));
    }
    if (s.nco_cos) {  // Signed8BitIQConverter.java:119-120: four rounded products, then -/+
        const int t = (int)(((long long)s.nco_idx + k) % s.nco_len);
        const float c = s.nco_cos[t], sn = s.nco_sin[t];
        re = __fsub_rn(__fmul_rn(r, c), __fmul_rn(q, sn));
        im = __fadd_rn(__fmul_rn(q, c), __fmul_rn(r, sn));
    } else {
        re = r;
        im = q;
    }
}

// ---- K4 (+K2) ---------------------------------------------------------------------------
struct ResampleArgs {
    StreamSrc src;
    const float *bank;  // [I][nt]
    int I, D, nt;
    long long rel;      // stream index (relative to this call's sample 0) of output 0
    int ph0;            // polyphase index of output 0
    long long nout;
    int tile;           // outputs per CTA
    int span_max;       // shared-memory capacity in samples
    float *out_re, *out_im;
};

template <int KIND, bool EXACT>
__global__ void __launch_bounds__(256) resample_kernel(const ResampleArgs a) {
    extern __shared__ float smem[];
    float *sre = smem, *sim = smem + a.span_max;
    const long long j0 = (long long)blockIdx.x * a.tile;
    long long j1 = j0 + a.tile;
    if (j1 > a.nout) j1 = a.nout;
    if (j0 >= j1) return;
    // stream window of this tile: [k_lo, k_hi]
    const long long k_first = a.rel + ((long long)a.ph0 + j0 * a.D) / a.I;
    const long long k_hi = a.rel + ((long long)a.ph0 + (j1 - 1) * a.D) / a.I;
    const long long k_lo = k_first - (a.nt - 1);
    const int span = (int)(k_hi - k_lo + 1);
    for (int s = threadIdx.x; s < span; s += blockDim.x) {
        float r, q;
        fetch<KIND>(a.src, k_lo + s, r, q);
        sre[s] = r;
        sim[s] = q;
    }
    __syncthreads();
    for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const long long T = (long long)a.ph0 + j * a.D;
        const int pos = (int)(a.rel + T / a.I - k_lo);  // newest sample of this output's window
        const float *taps = a.bank + (size_t)(T % a.I) * a.nt;
        float ar = 0.0f, ai = 0.0f;
        for (int t = 0; t < a.nt; t++) {  // RationalResampler.kt:124-131, oldest tap last
            const float h = __ldg(taps + t);
            ar = mac<EXACT>(ar, h, sre[pos - t]);
            ai = mac<EXACT>(ai, h, sim[pos - t]);
        }
        a.out_re[j] = ar;
        a.out_im[j] = ai;
    }
}

// ---- K5 ------------------------------------------------------------------------------------
struct FirArgs {
    StreamSrc src;
    const float *taps_re, *taps_im;  // taps_im != nullptr: complex taps
    int ntaps, dec;
    long long first;  // input index of output 0
    long long nout;
    int tile, span_max;
    int real_only;    // filterReal: imaginary channel neither read nor written
    float *out_re, *out_im;
};

template <bool CPLX, bool EXACT>
__global__ void __launch_bounds__(256) fir_kernel(const FirArgs a) {
    extern __shared__ float smem[];
    float *sre = smem, *sim = smem + a.span_max;
    float *str_ = smem + 2 * a.span_max, *sti = str_ + a.ntaps;
    for (int t = threadIdx.x; t < a.ntaps; t += blockDim.x) {
        str_[t] = a.taps_re[t];
        if (CPLX) sti[t] = a.taps_im[t];
    }
    const long long j0 = (long long)blockIdx.x * a.tile;
    long long j1 = j0 + a.tile;
    if (j1 > a.nout) j1 = a.nout;
    if (j0 >= j1) return;
    const long long k_lo = a.first + j0 * a.dec - (a.ntaps - 1);
    const long long k_hi = a.first + (j1 - 1) * a.dec;
    const int span = (int)(k_hi - k_lo + 1);
    for (int s = threadIdx.x; s < span; s += blockDim.x) {
        float r, q;
        fetch<3>(a.src, k_lo + s, r, q);
        sre[s] = r;
        sim[s] = q;
    }
    __syncthreads();
    for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const int pos = (int)(a.first + j * a.dec - k_lo);
        float ar = 0.0f, ai = 0.0f;
        if (CPLX) {  // ComplexFirFilter.java:147-151
            for (int t = 0; t < a.ntaps; t++) {
                const float tr = str_[t], ti = sti[t], xr = sre[pos - t], xi = sim[pos - t];
                if (EXACT) {
                    ar = __fadd_rn(ar, __fsub_rn(__fmul_rn(tr, xr), __fmul_rn(ti, xi)));
                    ai = __fadd_rn(ai, __fadd_rn(__fmul_rn(ti, xr), __fmul_rn(tr, xi)));
                } else {
                    ar += fmaf(tr, xr, -(ti * xi));
                    ai += fmaf(ti, xr, tr * xi);
                }
            }
        } else if (a.real_only) {  // FirFilter.kt:141-146
            for (int t = 0; t < a.ntaps; t++) ar = mac<EXACT>(ar, str_[t], sre[pos - t]);
        } else {  // FirFilter.kt:90-96
            for (int t = 0; t < a.ntaps; t++) {
                const float h = str_[t];
                ar = mac<EXACT>(ar, h, sre[pos - t]);
                ai = mac<EXACT>(ai, h, sim[pos - t]);
            }
        }
        a.out_re[j] = ar;
        if (!a.real_only) a.out_im[j] = ai;
    }
}

// new history = the last `hist` samples of (old history ++ consumed input)
template <int KIND>
__global__ void history_kernel(const StreamSrc src, long long consumed, float *new_re, float *new_im) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= src.hist) return;
    float r, q;
    fetch<KIND>(src, consumed - src.hist + h, r, q);
    new_re[h] = r;
    if (new_im) new_im[h] = q;
}

}  // namespace

static const int kSpanMax = 6144;  // samples staged per CTA: 2 * 6144 * 4 B = 48 KiB

static StreamSrc make_src(const StreamDesc &d) {
    StreamSrc s{};
    s.re = d.re;
    s.im = d.im;
    s.raw = d.raw;
    s.hist_re = d.hist_re;
    s.hist_im = d.hist_im;
    s.hist = d.hist;
    s.nco_cos = d.nco_cos;
    s.nco_sin = d.nco_sin;
    s.nco_len = d.nco_len;
    s.nco_idx = d.nco_idx;
    return s;
}

cudaError_t resample_launch(const StreamDesc &in, const float *bank, int I, int D, int nt, long long rel, int ph0,
                            long long nout, float *out_re, float *out_im, bool exact, cudaStream_t st) {
    if (nout <= 0) return cudaSuccess;
    if (nt + 2 > kSpanMax) return cudaErrorInvalidValue;
    ResampleArgs a{};
    a.src = make_src(in);
    a.bank = bank;
    a.I = I;
    a.D = D;
    a.nt = nt;
    a.rel = rel;
    a.ph0 = ph0;
    a.nout = nout;
    a.span_max = kSpanMax;
    // span of a tile <= floor((ph + (tile-1)*D)/I) + nt <= ((tile-1)*D + I-1)/I + nt
    long long tile = ((long long)(kSpanMax - nt - 2) * I) / D;
    if (tile < 1) return cudaErrorInvalidValue;  // D/I so large that one output's window overflows
    if (tile > 1024) tile = 1024;
    a.tile = (int)tile;
    a.out_re = out_re;
    a.out_im = out_im;
    const unsigned grid = (unsigned)((nout + tile - 1) / tile);
    const size_t smem = 2 * (size_t)kSpanMax * sizeof(float);
#define RFA_RS(KIND)                                                                              \
    do {                                                                                          \
        if (exact)                                                                                \
            resample_kernel<KIND, true><<<grid, 256, smem, st>>>(a);                              \
        else                                                                                      \
            resample_kernel<KIND, false><<<grid, 256, smem, st>>>(a);                             \
    } while (0)
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(resample_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    switch (in.kind) {
        case 0: RFA_RS(0); break;
        case 1: RFA_RS(1); break;
        case 2: RFA_RS(2); break;
        case 3: RFA_RS(3); break;
        default: return cudaErrorInvalidValue;
    }
#undef RFA_RS
    return cudaGetLastError();
}

cudaError_t fir_launch(const StreamDesc &in, const float *taps_re, const float *taps_im, int ntaps, int dec,
                       long long first, long long nout, bool real_only, float *out_re, float *out_im, bool exact,
                       cudaStream_t st) {
    if (nout <= 0) return cudaSuccess;
    if (ntaps + dec + 2 > kSpanMax || ntaps > 4096) return cudaErrorInvalidValue;
    FirArgs a{};
    a.src = make_src(in);
    a.taps_re = taps_re;
    a.taps_im = taps_im;
    a.ntaps = ntaps;
    a.dec = dec;
    a.first = first;
    a.nout = nout;
    a.span_max = kSpanMax;
    long long tile = (kSpanMax - ntaps) / dec;
    if (tile < 1) return cudaErrorInvalidValue;
    if (tile > 1024) tile = 1024;
    a.tile = (int)tile;
    a.real_only = real_only ? 1 : 0;
    a.out_re = out_re;
    a.out_im = out_im;
    const unsigned grid = (unsigned)((nout + tile - 1) / tile);
    const size_t smem = (2 * (size_t)kSpanMax + 2 * (size_t)ntaps) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        const int mx = (int)((2 * (size_t)kSpanMax + 2 * 4096) * sizeof(float));
        cudaFuncSetAttribute(fir_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        configured = true;
    }
    if (taps_im) {
        if (exact)
            fir_kernel<true, true><<<grid, 256, smem, st>>>(a);
        else
            fir_kernel<true, false><<<grid, 256, smem, st>>>(a);
    } else {
        if (exact)
            fir_kernel<false, true><<<grid, 256, smem, st>>>(a);
        else
            fir_kernel<false, false><<<grid, 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t history_launch(const StreamDesc &in, long long consumed, float *new_re, float *new_im,
                           cudaStream_t st) {
    if (in.hist <= 0) return cudaSuccess;
    const StreamSrc s = make_src(in);
    const unsigned grid = (unsigned)((in.hist + 127) / 128);
    switch (in.kind) {
        case 0: history_kernel<0><<<grid, 128, 0, st>>>(s, consumed, new_re, new_im); break;
        case 1: history_kernel<1><<<grid, 128, 0, st>>>(s, consumed, new_re, new_im); break;
        case 2: history_kernel<2><<<grid, 128, 0, st>>>(s, consumed, new_re, new_im); break;
        case 3: history_kernel<3><<<grid, 128, 0, st>>>(s, consumed, new_re, new_im); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace rfa
