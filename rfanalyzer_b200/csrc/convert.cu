// convert.cu -- K1 iq_convert and K2 iq_convert_mix.
//
//   K1 = IQConverter.fillPacketIntoSamplePacket   (A/source/Signed8BitIQConverter.java:80-99,
//        Unsigned8BitIQConverter.java:80-99, Signed16BitIQConverter.kt:89-124)
//   K2 = IQConverter.mixPacketIntoSamplePacket    (…java:102-131, …kt:126-181)
// Both are pure streaming kernels: 128-bit loads of the interleaved bytes, 128-bit stores of
// the planar floats, grid sized to the machine.  The reference's look-up tables are replaced by
// arithmetic that is exact for every code point (rfa_fft_core.cuh), and the mixer keeps the
// reference's rounding: each of the four products is rounded to float before the add/sub,
// exactly what its precomputed [t][byte] product tables hold.
#include <cuda_runtime.h>

#include "kernels.h"
#include "rfa_fft_core.cuh"

namespace rfa {
namespace {

template <int FMT>
__device__ __forceinline__ void decode(const void *iq, long long n, float &re, float &im) {
    if (FMT == FMT_S8) {
        uint16_t raw = ((const uint16_t *)iq)[n];
        re = conv_s8((int)(int8_t)(raw & 0xFF));
        im = conv_s8((int)(int8_t)(raw >> 8));
    } else if (FMT == FMT_U8) {
        uint16_t raw = ((const uint16_t *)iq)[n];
        re = conv_u8((int)(raw & 0xFF));
        im = conv_u8((int)(raw >> 8));
    } else {
        uint32_t raw = ((const uint32_t *)iq)[n];
        re = conv_s16((int)(int16_t)(raw & 0xFFFF));
        im = conv_s16((int)(int16_t)(raw >> 16));
    }
}

// SPT samples per 16-byte load: 8 for the 8-bit formats, 4 for s16
template <int FMT, bool MIX>
__global__ void __launch_bounds__(256) convert_kernel(const void *__restrict__ iq, long long nsamples,
                                                       float *__restrict__ re, float *__restrict__ im,
                                                       const float *__restrict__ cosT,
                                                       const float *__restrict__ sinT, int ncoLen, int ncoIdx,
                                                       bool vec_ok) {
    constexpr int SPT = (FMT == FMT_S16LE) ? 4 : 8;
    __shared__ float sc[512], ss[512];
    if (MIX) {
        for (int i = threadIdx.x; i < ncoLen; i += blockDim.x) {
            sc[i] = cosT[i];
            ss[i] = sinT[i];
        }
        __syncthreads();
    }
    const long long nvec = vec_ok ? nsamples / SPT : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 raw = __ldcs(((const uint4 *)iq) + v);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float r[SPT], q[SPT];
#pragma unroll
        for (int j = 0; j < SPT; j++) {
            if (FMT == FMT_S16LE) {
                r[j] = conv_s16((int)(int16_t)(w[j] & 0xFFFF));
                q[j] = conv_s16((int)(int16_t)(w[j] >> 16));
            } else {
                const uint32_t pair = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
                if (FMT == FMT_S8) {
                    r[j] = conv_s8((int)(int8_t)(pair & 0xFF));
                    q[j] = conv_s8((int)(int8_t)(pair >> 8));
                } else {
                    r[j] = conv_u8((int)(pair & 0xFF));
                    q[j] = conv_u8((int)(pair >> 8));
                }
            }
        }
        if (MIX) {
            int t = (int)(((long long)ncoIdx + v * SPT) % ncoLen);
#pragma unroll
            for (int j = 0; j < SPT; j++) {
                const float c = sc[t], s = ss[t];
                const float a = __fmul_rn(r[j], c), b = __fmul_rn(q[j], s);
                const float d = __fmul_rn(q[j], c), e = __fmul_rn(r[j], s);
                r[j] = __fsub_rn(a, b);
                q[j] = __fadd_rn(d, e);
                if (++t == ncoLen) t = 0;
            }
        }
#pragma unroll
        for (int j = 0; j < SPT; j += 4) {
            __stcs(((float4 *)re) + (v * SPT + j) / 4, make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]));
            __stcs(((float4 *)im) + (v * SPT + j) / 4, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
        }
    }
    // scalar tail (and the whole buffer when it is not 16-byte aligned)
    for (long long n = nvec * SPT + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < nsamples; n += stride) {
        float r, q;
        decode<FMT>(iq, n, r, q);
        if (MIX) {
            const int t = (int)(((long long)ncoIdx + n) % ncoLen);
            const float c = sc[t], s = ss[t];
            const float a = __fmul_rn(r, c), b = __fmul_rn(q, s);
            const float d = __fmul_rn(q, c), e = __fmul_rn(r, s);
            r = __fsub_rn(a, b);
            q = __fadd_rn(d, e);
        }
        re[n] = r;
        im[n] = q;
    }
}

template <int FMT, bool MIX>
cudaError_t launch(const void *iq, long long n, float *re, float *im, const float *c, const float *s, int len,
                   int idx, int num_sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    constexpr int SPT = (FMT == FMT_S16LE) ? 4 : 8;
    const bool vec_ok = (((uintptr_t)iq | (uintptr_t)re | (uintptr_t)im) & 15) == 0;
    long long work = vec_ok ? (n + SPT - 1) / SPT : n;
    long long blocks = (work + 255) / 256;
    long long cap = (long long)num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    convert_kernel<FMT, MIX><<<(unsigned)blocks, 256, 0, st>>>(iq, n, re, im, c, s, len, idx, vec_ok);
    return cudaGetLastError();
}

}  // namespace

cudaError_t convert_launch(int fmt, const void *iq, long long n, float *re, float *im, int num_sms,
                           cudaStream_t st) {
    switch (fmt) {
        case FMT_S8: return launch<FMT_S8, false>(iq, n, re, im, nullptr, nullptr, 0, 0, num_sms, st);
        case FMT_U8: return launch<FMT_U8, false>(iq, n, re, im, nullptr, nullptr, 0, 0, num_sms, st);
        case FMT_S16LE: return launch<FMT_S16LE, false>(iq, n, re, im, nullptr, nullptr, 0, 0, num_sms, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t mix_launch(int fmt, const void *iq, long long n, float *re, float *im, const float *cosT,
                       const float *sinT, int len, int idx, int num_sms, cudaStream_t st) {
    if (len < 1 || len > 512) return cudaErrorInvalidValue;
    switch (fmt) {
        case FMT_S8: return launch<FMT_S8, true>(iq, n, re, im, cosT, sinT, len, idx, num_sms, st);
        case FMT_U8: return launch<FMT_U8, true>(iq, n, re, im, cosT, sinT, len, idx, num_sms, st);
        case FMT_S16LE: return launch<FMT_S16LE, true>(iq, n, re, im, cosT, sinT, len, idx, num_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace rfa
