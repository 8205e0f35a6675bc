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
#include <stdlib.h>

#include "device_once.h"
#include "kernels.h"
#include "pdl.h"
#include "rfa_fft_core.cuh"
#include "spectrum_kernel.cuh"  // decode_point: magic-number sample conversion

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
    // the delay line after this call (the last src.hist samples of old history ++ `consumed` inputs): written by CTA 0 of
    // the tiled / stripe kernels before it starts on its tiles, instead of a history_kernel launch behind the resampler
    float *hist_new_re, *hist_new_im;
    long long consumed;
};

template <int KIND>
__device__ __forceinline__ void slide_history(const ResampleArgs &a) {
    if (blockIdx.x != 0 || !a.hist_new_re) return;
    for (int h = threadIdx.x; h < a.src.hist; h += blockDim.x) {
        float r, q;
        fetch<KIND>(a.src, a.consumed - a.src.hist + h, r, q);
        a.hist_new_re[h] = r;
        a.hist_new_im[h] = q;
    }
}

template <int KIND, bool EXACT>
__global__ void __launch_bounds__(256) resample_kernel(const ResampleArgs a) {
    pdl_trigger();  // the small kernels behind the resampler may be scheduled now (pdl.h); they wait for this grid to finish
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
    // T = ph0 + j*D decides window (T / I) and phase (T % I): one 64-bit division per CTA, 32-bit ones per output
    const long long T0 = (long long)a.ph0 + j0 * a.D;
    const long long q0 = T0 / a.I;
    const unsigned int r0 = (unsigned int)(T0 % a.I), uI = (unsigned int)a.I, uD = (unsigned int)a.D;
    const int base = (int)(a.rel + q0 - k_lo);
    for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const unsigned int u = r0 + (unsigned int)(j - j0) * uD;  // < I + tile*D < 2^31
        const int pos = base + (int)(u / uI);  // newest sample of this output's window
        const float *taps = a.bank + (size_t)(u % uI) * a.nt;
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

// ---- K4 (+K2), fast path (RFA_SUM_FMA) --------------------------------------------------------
// Same outputs as resample_kernel<KIND,false> up to the order of the float32 additions.
//  * staging: four input samples per thread and step (one 64/128-bit load of raw codes), NCO phase
//    advanced incrementally (no 64-bit modulo per sample), re/im interleaved in shared memory;
//  * dot products: G lanes share one output, lane g takes taps g, g+G, ... (neighbouring lanes read
//    neighbouring samples and taps: conflict-free, coalesced), one LDS.64 + one packed FFMA2 per tap
//    and sample, then a shuffle reduction.  G is chosen by the launcher so that a tile's outputs
//    times G fill the CTA (a 625/6 decimation yields only 54 outputs per staged span).
//  * the polyphase bank sits in shared memory when it fits.
struct ResampleFastArgs {
    ResampleArgs a;
    int G;          // lanes per output: 1, 2, 4, ... 32
    int bank_smem;  // bank floats copied to shared memory (0 = read through L1)
};

// one stream sample k >= 0: raw word first (so that several loads are in flight before any is converted)
template <int KIND>
__device__ __forceinline__ void load1(const StreamSrc &s, long long k, uint32_t &raw, float &im) {
    if (KIND == FMT_S8 || KIND == FMT_U8) {
        raw = (uint32_t)__ldg((const uint16_t *)s.raw + k);
    } else if (KIND == FMT_S16LE) {
        raw = __ldg((const uint32_t *)s.raw + k);
    } else {
        raw = __float_as_uint(s.re[k]);
        im = s.im ? s.im[k] : 0.0f;
    }
}
// ... then conversion + NCO mixing; t = NCO table index of the sample
template <int KIND>
__device__ __forceinline__ float2 convert1(const StreamSrc &s, uint32_t raw, float im, int t) {
    float r, q;
    if (KIND == FMT_S8) {
        r = conv_s8((int)(int8_t)(raw & 0xFF));
        q = conv_s8((int)(int8_t)(raw >> 8));
    } else if (KIND == FMT_U8) {
        r = conv_u8((int)(raw & 0xFF));
        q = conv_u8((int)(raw >> 8));
    } else if (KIND == FMT_S16LE) {
        r = conv_s16((int)(int16_t)(raw & 0xFFFF));
        q = conv_s16((int)(int16_t)(raw >> 16));
    } else {
        r = __uint_as_float(raw);
        q = im;
    }
    if (KIND != 3 && s.nco_cos) {  // Signed8BitIQConverter.java:119-120: four rounded products, then -/+
        const float c = __ldg(s.nco_cos + t), sn = __ldg(s.nco_sin + t);
        return make_float2(__fsub_rn(__fmul_rn(r, c), __fmul_rn(q, sn)), __fadd_rn(__fmul_rn(q, c), __fmul_rn(r, sn)));
    }
    return make_float2(r, q);
}
// Stage samples [k_al, k_al + span) of the stream into xs[0 .. span).  Consecutive lanes take consecutive
// samples: the raw loads, the NCO table reads and the 64-bit shared-memory stores are all stride-1 across a
// warp (four consecutive samples per thread cost a 4-way bank conflict on every store, four cache lines per
// table read and a wrap-around loop per sample).
template <int KIND>
__device__ __forceinline__ void stage_span(const StreamSrc &src, long long k_al, int span, float2 *xs) {
    const int nco_len = src.nco_len > 0 ? src.nco_len : 1;
    // NCO index of sample k_al + threadIdx.x, then advanced by blockDim.x per round
    long long t0 = ((long long)src.nco_idx + k_al + threadIdx.x) % nco_len;
    if (t0 < 0) t0 += nco_len;
    int t = (int)t0;
    const int sstep = (int)blockDim.x, tstep = sstep % nco_len;
    constexpr int U = 8;  // loads in flight per thread: the staging is a chain of DRAM round trips otherwise
    for (int s0 = threadIdx.x; s0 < span; s0 += U * sstep) {
        uint32_t raw[U];
        float rim[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int sidx = s0 + u * sstep;
            const long long k = k_al + sidx;
            rim[u] = 0.0f;
            raw[u] = 0u;
            if (sidx < span && k >= 0) load1<KIND>(src, k, raw[u], rim[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int sidx = s0 + u * sstep;
            const long long k = k_al + sidx;
            if (sidx < span) {
                if (k >= 0) {
                    xs[sidx] = convert1<KIND>(src, raw[u], rim[u], t);
                } else {  // history (and the zeros before it)
                    float r, q;
                    fetch<KIND>(src, k, r, q);
                    xs[sidx] = make_float2(r, q);
                }
            }
            t += tstep;
            if (t >= nco_len) t -= nco_len;
        }
    }
}
// Staging for the hot case of the tiled and stripe kernels: integer IQ, the whole span inside this call's input
// (k_al >= 0), NCO table in shared memory as (cos, sin) pairs PRE-SCALED by the format's power-of-two unit (scaling by a
// power of two commutes with rounding: same bits) and with entry nco_len repeating entry 0 (the second sample of a pair
// needs no wrap test).  A thread takes two consecutive samples per round -- one 64-bit (int16) or 32-bit (8-bit IQ)
// load, magic-number conversion to the integer code without I2F, three packed FP32 instructions for the mixer, one
// 128-bit conflict-free store -- about a dozen instructions per sample where stage_span spends sixty.
// `raw_s` != nullptr: the raw codes of the span already sit in shared memory (bulk copy), nothing waits on DRAM here.
template <int KIND>
__device__ __forceinline__ cf decode_code(uint32_t raw) {  // the sample as (I, Q) in units of one code step
    if (KIND == FMT_S8) {
        raw ^= 0x8080u;
        return cadd(cf{magic_byte0(raw), magic_byte1(raw)}, cf{-8388736.0f, -8388736.0f});
    } else if (KIND == FMT_U8) {
        return cadd(cadd(cf{magic_byte0(raw), magic_byte1(raw)}, cf{-8388608.0f, -8388608.0f}), cf{-127.4f, -127.4f});
    } else {
        raw ^= 0x80008000u;
        return cadd(cf{magic_half0(raw), magic_half1(raw)}, cf{-8421376.0f, -8421376.0f});
    }
}
template <int KIND>
__device__ __forceinline__ void stage_span_pairs(const StreamSrc &src, long long k_al, int span, float2 *xs, const float2 *s_nco,
                                                 int t_first /* NCO index of sample k_al */, const void *raw_s = nullptr) {
    const int nco_len = src.nco_len > 0 ? src.nco_len : 1;
    const int npairs = span >> 1;  // an odd last sample is left to the caller
    int t = (t_first + 2 * (int)threadIdx.x) % nco_len;
    const int tstep = (2 * (int)blockDim.x) % nco_len;
    const bool mixing = src.nco_cos != nullptr;
    constexpr float unit = KIND == FMT_S16LE ? (1.0f / 32768.0f) : 0.0078125f;
    constexpr int U = 8;  // loads in flight per thread: a chain of DRAM round trips otherwise
    for (int pr0 = threadIdx.x; pr0 < npairs; pr0 += U * blockDim.x) {
        uint32_t r0[U], r1[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int pr = pr0 + u * (int)blockDim.x;
            r0[u] = r1[u] = 0u;
            if (pr < npairs) {
                if (KIND == FMT_S16LE) {
                    const uint2 v = raw_s ? reinterpret_cast<const uint2 *>(raw_s)[pr]
                                          : __ldg(reinterpret_cast<const uint2 *>((const uint32_t *)src.raw + k_al) + pr);
                    r0[u] = v.x;
                    r1[u] = v.y;
                } else {
                    const uint32_t v = raw_s ? reinterpret_cast<const uint32_t *>(raw_s)[pr]
                                             : __ldg(reinterpret_cast<const uint32_t *>((const uint16_t *)src.raw + k_al) + pr);
                    r0[u] = v & 0xFFFFu;
                    r1[u] = v >> 16;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int pr = pr0 + u * (int)blockDim.x;
            cf a = decode_code<KIND>(r0[u]), b = decode_code<KIND>(r1[u]);
            if (mixing) {
                const float2 c0 = s_nco[t], c1 = s_nco[t + 1];  // (cos, sin) * unit
                // re = r*c - q*s, im = q*c + r*s, every product rounded on its own (Signed8BitIQConverter.java:119-120):
                // the same bits as fetch<KIND> / stage_span produce, whichever tile or call a sample is staged in
                // (packed products, SCALAR sums: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 in spite of the
                // explicit rounding modifiers -- measured, build/dbg/stage_dbg.cu -- and only the scalar intrinsics are
                // guaranteed to stay separate)
                const cf p0 = cscale(a, c0.x), q0 = cscale(cf{a.y, a.x}, c0.y);
                a = cf{__fsub_rn(p0.x, q0.x), __fadd_rn(p0.y, q0.y)};
                const cf p1 = cscale(b, c1.x), q1 = cscale(cf{b.y, b.x}, c1.y);
                b = cf{__fsub_rn(p1.x, q1.x), __fadd_rn(p1.y, q1.y)};
                t += tstep;
                if (t >= nco_len) t -= nco_len;
            } else {
                a = cscale(a, unit);
                b = cscale(b, unit);
            }
            if (pr < npairs) *reinterpret_cast<float4 *>(xs + 2 * pr) = make_float4(a.x, a.y, b.x, b.y);
        }
    }
}

// the pre-scaled (cos, sin) table stage_span_pairs reads: nco_len + 1 entries
template <int KIND>
__device__ __forceinline__ void fill_nco_pairs(const StreamSrc &src, float2 *s_nco) {
    const float unit = KIND == FMT_S16LE ? (1.0f / 32768.0f) : ((KIND == FMT_S8 || KIND == FMT_U8) ? 0.0078125f : 1.0f);
    for (int i = threadIdx.x; i <= src.nco_len; i += blockDim.x) {
        const int k = i == src.nco_len ? 0 : i;
        s_nco[i] = make_float2(__ldg(src.nco_cos + k) * unit, __ldg(src.nco_sin + k) * unit);
    }
}

template <int KIND, bool BANK_SMEM, int G>
__global__ void __launch_bounds__(256) resample_fast_kernel(const ResampleFastArgs fa) {
    pdl_trigger();  // the small kernels behind the resampler may be scheduled now (pdl.h); they wait for this grid to finish
    const ResampleArgs &a = fa.a;
    extern __shared__ float2 xs[];  // [span_max + 8] samples, then the bank
    float *sbank = reinterpret_cast<float *>(xs + a.span_max + 8);
    const long long j0 = (long long)blockIdx.x * a.tile;
    long long j1 = j0 + a.tile;
    if (j1 > a.nout) j1 = a.nout;
    if (j0 >= j1) return;
    const long long k_first = a.rel + ((long long)a.ph0 + j0 * a.D) / a.I;
    const long long k_hi = a.rel + ((long long)a.ph0 + (j1 - 1) * a.D) / a.I;
    const long long k_lo = k_first - (a.nt - 1);
    // stage [k_al, k_hi], k_al = k_lo rounded down to a multiple of 4 (floor also for negatives)
    const long long k_al = k_lo - (((k_lo % 4) + 4) % 4);
    const int span = (int)(k_hi - k_al + 1);
    if (BANK_SMEM)
        for (int i = threadIdx.x; i < fa.bank_smem; i += blockDim.x) sbank[i] = a.bank[i];
    stage_span<KIND>(a.src, k_al, span, xs);
    __syncthreads();
    const int g = threadIdx.x & (G - 1), grp = threadIdx.x / G, ngrp = blockDim.x / G;
    // T = ph0 + j*D decides window (T / I) and phase (T % I): one 64-bit division per CTA, 32-bit ones per output
    const long long T0 = (long long)a.ph0 + j0 * a.D;
    const long long q0 = T0 / a.I;
    const unsigned int r0 = (unsigned int)(T0 % a.I), uI = (unsigned int)a.I, uD = (unsigned int)a.D;
    const int base = (int)(a.rel + q0 - k_al);
    for (long long jb = j0; jb < j1; jb += ngrp) {  // warp-uniform trip count (the shuffles below)
        const long long j = jb + grp;
        const bool valid = j < j1;
        const unsigned int u = r0 + (valid ? (unsigned int)(j - j0) : 0u) * uD;  // < I + tile*D < 2^31
        const int pos = base + (int)(u / uI);
        const size_t tap0 = (size_t)(u % uI) * a.nt;
        const float *taps = BANK_SMEM ? sbank + tap0 : a.bank + tap0;  // two address spaces, two code paths
        const float2 *x = xs + pos;
        cf acc = cf{0.0f, 0.0f}, acc2 = cf{0.0f, 0.0f};
        int t = valid ? g : a.nt;
        for (; t + 3 * G < a.nt; t += 4 * G) {  // four taps per step on two independent chains
            float2 xv[4];
            float hv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                xv[k] = x[-(t + k * G)];
                hv[k] = BANK_SMEM ? taps[t + k * G] : __ldg(taps + t + k * G);
            }
            acc = caxpy(hv[0], cf{xv[0].x, xv[0].y}, acc);
            acc2 = caxpy(hv[1], cf{xv[1].x, xv[1].y}, acc2);
            acc = caxpy(hv[2], cf{xv[2].x, xv[2].y}, acc);
            acc2 = caxpy(hv[3], cf{xv[3].x, xv[3].y}, acc2);
        }
        for (; t < a.nt; t += G) {
            const float2 x0 = x[-t];
            acc = caxpy(BANK_SMEM ? taps[t] : __ldg(taps + t), cf{x0.x, x0.y}, acc);
        }
        acc = cadd(acc, acc2);
        for (int d = G >> 1; d > 0; d >>= 1) {
            acc.x += __shfl_xor_sync(0xFFFFFFFFu, acc.x, d);
            acc.y += __shfl_xor_sync(0xFFFFFFFFu, acc.y, d);
        }
        if (g == 0 && valid) {
            a.out_re[j] = acc.x;
            a.out_im[j] = acc.y;
        }
    }
}

// ---- K4 (+K2), register-tiled path for moderate decimation (taps per phase > D) -----------------
// Outputs of one polyphase share their taps and their windows slide by exactly D samples, so with
// t = a*D + b the sum  y[m] = sum_t h[t] x[pos0 + m*D - t]  becomes, for every b, a short convolution
// over (a, m) of h_b[a] = h[a*D + b] with s_b[n] = x[pos0 - b + n*D].  One thread owns M consecutive
// outputs of one phase: per b it loads M + A - 1 samples and A taps and issues M*A packed FFMA2 --
// about three times fewer shared-memory loads per FMA than one load pair per tap, and with M*D odd
// the sample loads of a half-warp fall into distinct banks.  The taps are transposed to [phase][b][a]
// in shared memory (zero-padded to AP per b) so that a thread reads them with 128-bit broadcasts.
struct ResampleTiledArgs {
    ResampleArgs a;  // a.tile = I * M * B outputs per CTA
    int B;           // blocks of M same-phase outputs per phase and CTA
    int A;           // ceil(nt / D) <= AMAX
    int AP;          // A rounded up to a multiple of 4
    int nco_pairs;   // room for the pre-scaled (cos, sin) table behind the bank (0: none, stage_span does everything)
};

template <int KIND, int M, int AMAX>
__global__ void __launch_bounds__(256) resample_tiled_kernel(const ResampleTiledArgs ta) {
    pdl_trigger();  // the small kernels behind the resampler may be scheduled now (pdl.h); they wait for this grid to finish
    const ResampleArgs &a = ta.a;
    slide_history<KIND>(a);
    constexpr int AP = (AMAX + 3) & ~3;
    extern __shared__ float2 xs[];  // [span_max + 8] samples, then the transposed bank [I][D][AP]
    float *hT = reinterpret_cast<float *>(xs + a.span_max + 8);
    const long long j0 = (long long)blockIdx.x * a.tile;
    if (j0 >= a.nout) return;
    const int nq = ta.B * M;  // outputs per phase in this tile
    // stream window of the tile: the newest sample belongs to the last output of the residue with the
    // largest offset, the oldest to tap index A*D - 1 of output j0
    const long long k_first = a.rel + ((long long)a.ph0 + j0 * a.D) / a.I;
    long long k_hi = a.rel + ((long long)a.ph0 + (j0 + a.I - 1) * a.D) / a.I + (long long)(nq - 1) * a.D;
    const long long k_end = a.rel + ((long long)a.ph0 + (a.nout - 1) * a.D) / a.I;  // newest sample any output needs
    if (k_hi > k_end) k_hi = k_end;  // last tile: never read past the input
    const long long k_lo = k_first - ((long long)AMAX * a.D - 1);  // taps a >= A are zero, their samples just have to exist
    const long long k_al = k_lo - (((k_lo % 4) + 4) % 4);
    const int span = (int)(k_hi - k_al + 1);
    for (int pb = threadIdx.x; pb < a.I * a.D; pb += blockDim.x) {  // one (phase, b) row of AP taps per step
        const int b = pb % a.D, p = pb / a.D;
#pragma unroll
        for (int aa = 0; aa < AP; aa++) {
            const int t = aa * a.D + b;
            hT[pb * AP + aa] = (aa < ta.A && t < a.nt) ? __ldg(a.bank + (size_t)p * a.nt + t) : 0.0f;
        }
    }
    constexpr bool INTFMT = KIND == FMT_S8 || KIND == FMT_U8 || KIND == FMT_S16LE;
    const bool mixing = a.src.nco_cos != nullptr;
    if (INTFMT && k_al >= 0 && ((size_t)a.src.raw & 7) == 0 && (!mixing || ta.nco_pairs > a.src.nco_len)) {
        float2 *s_nco = reinterpret_cast<float2 *>(hT + (size_t)a.I * a.D * AP);
        if (mixing) {
            fill_nco_pairs<KIND>(a.src, s_nco);
            __syncthreads();
        }
        const int t0 = (int)(((long long)a.src.nco_idx + k_al) % (a.src.nco_len > 0 ? a.src.nco_len : 1));
        stage_span_pairs<KIND>(a.src, k_al, span, xs, s_nco, t0);
        if ((span & 1) && threadIdx.x == 0) {
            float r, q;
            fetch<KIND>(a.src, k_al + span - 1, r, q);
            xs[span - 1] = make_float2(r, q);
        }
    } else {
        stage_span<KIND>(a.src, k_al, span, xs);
    }
    __syncthreads();
    for (int task = threadIdx.x; task < a.I * ta.B; task += blockDim.x) {
        const int r = task / ta.B, qb = task - r * ta.B;
        const long long T = (long long)a.ph0 + (j0 + r) * a.D;
        const int phase = (int)(T % a.I);
        const int pos0 = (int)(a.rel + T / a.I - k_al) + qb * M * a.D;
        cf acc[M];
#pragma unroll
        for (int m = 0; m < M; m++) acc[m] = cf{0.0f, 0.0f};
        const float *hp = hT + (size_t)phase * a.D * AP;
#pragma unroll 5  // sample and tap addresses become immediates inside the unrolled body (15 of 121 instructions per b were address updates)
        for (int b = 0; b < a.D; b++) {
            float h[AP];
#pragma unroll
            for (int v = 0; v < AP / 4; v++) {
                const float4 w = *reinterpret_cast<const float4 *>(hp + b * AP + 4 * v);
                h[4 * v] = w.x, h[4 * v + 1] = w.y, h[4 * v + 2] = w.z, h[4 * v + 3] = w.w;
            }
            const float2 *sp = xs + pos0 - b;
            cf sv[M + AMAX - 1];  // sv[i] = s_b[i - (AMAX - 1)]
#pragma unroll
            for (int i = 0; i < M + AMAX - 1; i++) {
                const float2 x = sp[(i - (AMAX - 1)) * a.D];
                sv[i] = cf{x.x, x.y};
            }
#pragma unroll
            for (int aa = 0; aa < AMAX; aa++)
#pragma unroll
                for (int m = 0; m < M; m++) acc[m] = caxpy(h[aa], sv[m - aa + AMAX - 1], acc[m]);
        }
#pragma unroll
        for (int m = 0; m < M; m++) {
            const long long j = j0 + (long long)(qb * M + m) * a.I + r;
            if (j < a.nout) {
                a.out_re[j] = acc[m].x;
                a.out_im[j] = acc[m].y;
            }
        }
    }
}


// ---- K4 (+K2), stripe path for large decimation (RFA_SUM_FMA) -----------------------------------------------------
// Airspy 10 Msps -> 96 kHz is I/D = 6/625 with 501 taps per phase: every input sample is used by 4.8 outputs, yet a
// load pair per multiply-add (resample_fast_kernel) makes the shared-memory pipe, not the 64 MB of input, the bound.
// Two facts of the polyphase structure give register reuse:
//   * outputs j and j + I have the SAME phase (taps) and windows exactly D samples apart, so one tap serves TP
//     periods:  acc[pi] += h * x[n + pi*D];
//   * neighbouring phases of one period look at almost the same samples (windows ~D/I apart), so one sample serves
//     PH phases:  acc[s] += h_s[P_s - n] * x[n]  (a tap index outside 0 .. nt-1 is a zero tap).
// A warp owns PH phases x TP periods; its lanes stripe over the union of those windows (neighbouring lanes read
// neighbouring samples and taps: conflict-free), each lane keeps PH*TP packed accumulators -- TP + PH loads per PH*TP
// packed multiply-adds -- and the 32 partial sums per output meet through shared memory.  A CTA stages TPC periods of
// decoded, NCO-mixed samples once (stage_span) and its warps share them.
struct ResampleStripeArgs {
    ResampleArgs a;  // a.tile = I * TPC outputs per CTA
    int TPC;         // periods per CTA (a multiple of TP)
    int pad;         // zero taps on either side of a bank row in shared memory
    int raw_bytes;   // shared-memory room for the next tile's raw codes (0: no bulk-copy staging)
    int acc_pairs;   // accumulators in front of the samples: one float pair per output AND per warp that shares a task's sample range
    int part_pairs;  // I * TPC rounded up to an even count: the accumulators of one such warp ("part")
    int nco_pairs;   // room for the (cos, sin) table
};

#ifdef RFA_STRIPE_TRACE
__device__ unsigned long long g_stripe_trace[8];
#define STRIPE_STAMP(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_stripe_trace[k], (unsigned long long)(now_ - t_prev_)); t_prev_ = now_; } } while (0)
#else
#define STRIPE_STAMP(k) do {} while (0)
#endif
template <int KIND, int PH, int TP, bool DODD>
__device__ __forceinline__ void resample_stripe_body(const ResampleStripeArgs &sa);
template <int KIND, int PH, int TP, bool DODD>
__global__ void __launch_bounds__(512) resample_stripe_kernel(const ResampleStripeArgs sa) {
    resample_stripe_body<KIND, PH, TP, DODD>(sa);
}
// the same body for two CTAs of TWELVE warps per SM (<= 85 registers; the default): a tile of I = 6 phases is six warp
// tasks, split in two over the sample range that is twelve -- no warp idles at the tile's barrier, which was the
// eight-warp kernel's top stall (ncu: 1.7 cycles per issue).  nFM 221 -> 209 us per 2^26 samples.
template <int KIND, int PH, int TP, bool DODD>
__global__ void __launch_bounds__(384, 2) resample_stripe12_kernel(const ResampleStripeArgs sa) {
    resample_stripe_body<KIND, PH, TP, DODD>(sa);
}
template <int KIND, int PH, int TP, bool DODD>
__device__ __forceinline__ void resample_stripe_body(const ResampleStripeArgs &sa) {
    pdl_trigger();  // the small kernels behind the resampler may be scheduled now (pdl.h); they wait for this grid to finish
#ifdef RFA_STRIPE_TRACE
    long long t_prev_ = clock64();
#endif
    const ResampleArgs &a = sa.a;
    slide_history<KIND>(a);
    // shared memory: [2 + span_max + 8] samples (two entries of slack in front: a lane may look one sample back), the
    // NCO table [512], the bank as zero-padded rows [I][pad + nt + pad], reduction scratch, slot table
    extern __shared__ float2 smem_stripe[];
    float2 *xs = smem_stripe + sa.acc_pairs + 2;  // [acc_pairs] output accumulators first (16-byte multiple), then two entries of slack
    float2 *s_nco = xs + a.span_max + 8;
    float *sbank = reinterpret_cast<float *>(s_nco + sa.nco_pairs);
    const int RS = a.nt + 2 * sa.pad;
    int *s_slot = reinterpret_cast<int *>(sbank + (size_t)a.I * RS);  // [I][2]: newest-sample position, tap row
    // raw IQ codes of the NEXT tile: one bulk copy (TMA) issued before this tile's dot products, landed when they end
    unsigned char *s_raw = reinterpret_cast<unsigned char *>(s_slot + 2 * a.I + 2);
    s_raw += (16 - ((size_t)s_raw & 15)) & 15;
    __shared__ int s_t0;
    __shared__ __align__(8) unsigned long long s_mbar;
    constexpr int BPS = KIND == FMT_S16LE ? 4 : 2;
    constexpr bool INTFMT = KIND == FMT_S8 || KIND == FMT_U8 || KIND == FMT_S16LE;
    const bool tma_ok = INTFMT && sa.raw_bytes > 0 && ((size_t)a.src.raw & 15) == 0;
    if (tma_ok && threadIdx.x == 0) mbar_init(&s_mbar);
    unsigned int waits = 0;  // completed bulk copies so far (CTA-uniform): the mbarrier's phase parity
    // geometry of a tile: first staged sample, samples, and how many of them a bulk copy may fetch
    // A tile is addressed by (qT, rT) = (T0 / I, T0 % I) of its first output's polyphase time T0 = ph0 + j0 * D; the CTA
    // steps from one of its tiles to the next by a constant (dq, dr), so the tile loop carries no 64-bit division.
    const long long k_end = a.rel + ((long long)a.ph0 + (a.nout - 1) * a.D) / a.I;  // newest sample any output needs
    auto tile_geom = [&](long long qT, int rT, long long *k_al_out, int *span_out, int *tma_samples) {
        const long long k_first = a.rel + qT;
        long long k_hi = k_first + (rT + (a.I - 1) * a.D) / a.I + (long long)(sa.TPC - 1) * a.D;
        if (k_hi > k_end) k_hi = k_end;  // last tile: never read past the input
        const long long k_lo = k_first - (a.nt - 1);
        constexpr int AL = 16 / BPS;  // a bulk copy starts on a 16-byte boundary of the input
        const long long k_al = k_lo & ~(long long)(AL - 1);
        const int span = (int)(k_hi - k_al + 1);
        *k_al_out = k_al;
        *span_out = span;
        // whole 16-byte units only (the copy must not run past the newest sample the call owns), whole pairs of samples
        int ts = (tma_ok && k_al >= 0) ? (int)(((long long)span * BPS) & ~15LL) / BPS : 0;
        if ((long long)ts * BPS > sa.raw_bytes) ts = 0;
        *tma_samples = ts;
    };
    long long qT;
    int rT;
    {
        const long long T0first = (long long)a.ph0 + (long long)blockIdx.x * a.tile * a.D;
        qT = T0first / a.I;
        rT = (int)(T0first % a.I);
    }
    const long long dT = (long long)gridDim.x * a.tile * a.D;
    const long long dq = dT / a.I;
    const int dr = (int)(dT % a.I);
    // once per CTA: the polyphase bank (zero taps around every row, so that the inner loop needs no range test) and
    // the NCO table; every load is issued before the first one is consumed -- a load/store pair per iteration is a
    // chain of L2 round trips.  Then the CTA walks tiles blockIdx.x, +gridDim.x, ...
    {
        const int nb = a.I * RS;
        constexpr int UB = 8;
        for (int i0 = threadIdx.x; i0 < nb; i0 += UB * blockDim.x) {
            float v[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int i = i0 + u * (int)blockDim.x, row = i / RS, t = i - row * RS - sa.pad;
                v[u] = (i < nb && t >= 0 && t < a.nt) ? __ldg(a.bank + (size_t)row * a.nt + t) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < UB; u++)
                if (i0 + u * (int)blockDim.x < nb) sbank[i0 + u * blockDim.x] = v[u];
        }
        if (a.src.nco_cos)
            fill_nco_pairs<KIND>(a.src, s_nco);
        if (threadIdx.x < 2) smem_stripe[sa.acc_pairs + threadIdx.x] = make_float2(0.0f, 0.0f);
    }
    const long long ntiles = (a.nout + a.tile - 1) / a.tile;
    __syncthreads();  // mbarrier initialised
    STRIPE_STAMP(0);  // prologue
    if (threadIdx.x == 0 && (long long)blockIdx.x < ntiles) {  // the first tile's raw codes start moving right away
        long long k0;
        int sp0, ts0;
        tile_geom(qT, rT, &k0, &sp0, &ts0);
        if (ts0) tma_load_1d(s_raw, (const char *)a.src.raw + k0 * BPS, (uint32_t)(ts0 * BPS), &s_mbar);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int groups = (a.I + PH - 1) / PH, blocks = sa.TPC / TP;
    static_assert(PH * TP * 2 <= 32, "one lane per reduced value");
    // A task = PH phases x TP periods; when a tile has fewer tasks than the CTA has warps, the sample range of a task is
    // split over `nsplit` warps and the partial sums meet in shared memory (s_acc, one float pair per output).
    const int ntasks = groups * blocks;
    const int nsplit = ntasks < nwarps ? nwarps / ntasks : 1;
    float *s_acc = reinterpret_cast<float *>(smem_stripe);  // [2 * I * TPC] -- lives in front of xs (see the launcher)
    const int task_w = warp / nsplit, part_w = warp - task_w * nsplit;
    const int s0_w = (task_w % groups) * PH, pi0_w = (task_w / groups) * TP;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long j0 = tile * a.tile;
    long long k_al;
    int span, tma_samples;
    tile_geom(qT, rT, &k_al, &span, &tma_samples);
    // the CTA's next tile
    long long qN = qT + dq;
    int rN = rT + dr;
    if (rN >= a.I) {
        rN -= a.I;
        qN++;
    }
    __syncthreads();  // the previous tile's warps are done with xs, the slot table and the output accumulators
    for (int i = threadIdx.x; i < sa.acc_pairs; i += blockDim.x) smem_stripe[i] = make_float2(0.0f, 0.0f);
    if ((int)threadIdx.x < a.I) {  // phase slot s of the tile: T = T0 + s * D
        const int e = rT + (int)threadIdx.x * a.D;  // < I + I * D: 32-bit
        s_slot[2 * threadIdx.x] = (int)(a.rel + qT + e / a.I - k_al);
        s_slot[2 * threadIdx.x + 1] = (e % a.I) * RS + sa.pad;
    }
    const bool fast = INTFMT && k_al >= 0 && ((size_t)a.src.raw & 7) == 0;
    if (fast) {
        if (threadIdx.x == 0) s_t0 = (int)(((long long)a.src.nco_idx + k_al) % (a.src.nco_len > 0 ? a.src.nco_len : 1));
        __syncthreads();
        STRIPE_STAMP(1);  // top-of-tile barriers and slot table
        if (tma_samples) {  // decode out of the bulk copy, then the few samples it could not carry
            mbar_wait(&s_mbar, waits & 1u);
            waits++;
            STRIPE_STAMP(2);  // wait for the bulk copy
            stage_span_pairs<KIND>(a.src, k_al, tma_samples, xs, s_nco, s_t0, s_raw);
            for (int i = tma_samples + threadIdx.x; i < span; i += blockDim.x) {
                float r, q;
                fetch<KIND>(a.src, k_al + i, r, q);
                xs[i] = make_float2(r, q);
            }
        } else {
            stage_span_pairs<KIND>(a.src, k_al, span, xs, s_nco, s_t0);
            if ((span & 1) && threadIdx.x == 0) {
                float r, q;
                fetch<KIND>(a.src, k_al + span - 1, r, q);
                xs[span - 1] = make_float2(r, q);
            }
        }
    } else {
        stage_span<KIND>(a.src, k_al, span, xs);
    }
    // samples past the input (periods of the last tile that no output uses): defined values, never NaN
    for (int i = span + threadIdx.x; i < a.span_max + 8; i += blockDim.x) xs[i] = make_float2(0.0f, 0.0f);
    STRIPE_STAMP(3);  // decode (thread 0's share)
    __syncthreads();
    STRIPE_STAMP(4);  // barrier after decode
    if (threadIdx.x == 0 && tile + gridDim.x < ntiles) {  // the raw buffer is free: fetch the next tile under the dot products
        long long kn;
        int spn, tsn;
        tile_geom(qN, rN, &kn, &spn, &tsn);
        if (tsn) {
            tma_load_1d(s_raw, (const char *)a.src.raw + kn * BPS, (uint32_t)(tsn * BPS), &s_mbar);
        } else if (INTFMT && kn >= 0) {
            // no shared-memory room for the raw codes (two CTAs per SM): at least pull the next tile's span into L2
            // while this tile's dot products run, so that its staging loads wait a third as long (whole 16-byte units
            // inside the input only)
            const size_t lo = ((size_t)a.src.raw + (size_t)kn * BPS + 15) & ~(size_t)15;
            const size_t hi = ((size_t)a.src.raw + (size_t)(kn + spn) * BPS) & ~(size_t)15;
            if (hi > lo)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
        }
    }
    for (int sub = warp; sub < ntasks * nsplit; sub += nwarps) {
        // (the warp's first sub-task is the same in every tile: its divisions were done before the tile loop)
        const int task = sub == warp ? task_w : sub / nsplit, part = sub == warp ? part_w : sub - task * nsplit;
        const int s0 = sub == warp ? s0_w : (task % groups) * PH, pi0 = sub == warp ? pi0_w : (task / groups) * TP;
        // per phase slot: position of its newest sample in xs (period pi0) and its (padded) tap row
        int base[PH];
        const float *tapb[PH];
#pragma unroll
        for (int s = 0; s < PH; s++) {
            const int slot = s0 + s < a.I ? s0 + s : a.I - 1;  // a short last group repeats its last slot (not written)
            base[s] = s_slot[2 * slot] + pi0 * a.D;
            tapb[s] = sbank + s_slot[2 * slot + 1] + base[s];  // tap of sample n: tapb[s][-n], zero outside 0 .. nt-1
        }
        cf acc[PH][TP];
#pragma unroll
        for (int s = 0; s < PH; s++)
#pragma unroll
            for (int q = 0; q < TP; q++) acc[s][q] = cf{0.0f, 0.0f};
        // A lane takes the sample PAIR (n, n+1), n even, of every period with one 128-bit load; where q*D is odd the
        // aligned pair is (n-1, n) instead -- every sample of every window is still visited exactly once.
        const int n_first = (base[0] - (a.nt - 1)) & ~1, n_last = base[PH - 1] + 1;
        const int chunks = (n_last - n_first) / 64 + 1;              // 64 samples per warp step
        int c_lo, c_hi;  // this warp's share of the chunks
        if (nsplit == 1) {
            c_lo = 0;
            c_hi = chunks;
        } else if (nsplit == 2) {
            c_lo = part ? chunks >> 1 : 0;
            c_hi = part ? chunks : chunks >> 1;
        } else {
            c_lo = chunks * part / nsplit;
            c_hi = chunks * (part + 1) / nsplit;
        }
        const float2 *xq = xs;
        for (int n = n_first + 64 * c_lo + 2 * lane; n < n_first + 64 * c_hi; n += 64) {
            float hm[PH], h0[PH], h1[PH];
#pragma unroll
            for (int s = 0; s < PH; s++) {
                h0[s] = tapb[s][-n];
                h1[s] = tapb[s][-n - 1];
                hm[s] = DODD ? tapb[s][-n + 1] : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < TP; q++) {
                const bool odd = DODD && (q & 1);
                const float4 v = *reinterpret_cast<const float4 *>(xq + n + q * a.D - (odd ? 1 : 0));
                const cf xa = cf{v.x, v.y}, xb = cf{v.z, v.w};
#pragma unroll
                for (int s = 0; s < PH; s++) {
                    acc[s][q] = caxpy(odd ? hm[s] : h0[s], xa, acc[s][q]);
                    acc[s][q] = caxpy(odd ? h0[s] : h1[s], xb, acc[s][q]);
                }
            }
        }
        // 32 lanes x 32 values -> lane v holds the total of value v: at every step half of the values a lane still
        // carries go to its partner (31 shuffles in all instead of 5 per value, no shared memory)
        float vals[32];
#pragma unroll
        for (int s = 0; s < PH; s++)
#pragma unroll
            for (int q = 0; q < TP; q++) {
                vals[(s * TP + q) * 2] = acc[s][q].x;
                vals[(s * TP + q) * 2 + 1] = acc[s][q].y;
            }
#pragma unroll
        for (int v = PH * TP * 2; v < 32; v++) vals[v] = 0.0f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; i++) {
                const float send = upper ? vals[i] : vals[i + off];
                const float keep = upper ? vals[i + off] : vals[i];
                vals[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
            }
        }
        if (lane < PH * TP * 2) {
            const int v = lane;
            const int s = (v >> 1) / TP, q = (v >> 1) % TP;
            // every part keeps its own copy (no atomics: the parts are added in a fixed order below, so the result does not
            // depend on which warp arrives first)
            if (s0 + s < a.I) s_acc[(size_t)part * 2 * sa.part_pairs + (((pi0 + q) * a.I) + s0 + s) * 2 + (v & 1)] = vals[0];
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.tile; i += blockDim.x) {  // tile outputs: period-major, phase slot, (re, im)
        const long long j = j0 + (i >> 1);
        if (j < a.nout) {
            float v = s_acc[i];
            for (int part = 1; part < nsplit; part++) v += s_acc[(size_t)part * 2 * sa.part_pairs + i];
            if (i & 1)
                a.out_im[j] = v;
            else
                a.out_re[j] = v;
        }
    }
    STRIPE_STAMP(5);  // dot products (warp 0's tasks)
    qT = qN;
    rT = rN;
  }  // tiles
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
    pdl_enter();
    extern __shared__ float smem[];
    // RFA_SUM_EXACT: planar samples and taps, every product and sum rounded on its own, in the reference's order.
    // RFA_SUM_FMA: samples and taps INTERLEAVED as (re, im) pairs in the same shared memory -- one 64-bit load each per
    // tap and one (real taps) or two (complex taps) packed multiply-adds instead of four loads and four scalar operations.
    float *sre = smem, *sim = smem + a.span_max;
    float *str_ = smem + 2 * a.span_max, *sti = str_ + a.ntaps;
    float2 *xs = reinterpret_cast<float2 *>(smem), *st = reinterpret_cast<float2 *>(smem + 2 * a.span_max);
    for (int t = threadIdx.x; t < a.ntaps; t += blockDim.x) {
        if (EXACT) {
            str_[t] = a.taps_re[t];
            if (CPLX) sti[t] = a.taps_im[t];
        } else {
            st[t] = make_float2(a.taps_re[t], CPLX ? a.taps_im[t] : 0.0f);
        }
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
        if (EXACT) {
            sre[s] = r;
            sim[s] = q;
        } else {
            xs[s] = make_float2(r, q);
        }
    }
    __syncthreads();
    for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const int pos = (int)(a.first + j * a.dec - k_lo);
        float ar = 0.0f, ai = 0.0f;
        if constexpr (!EXACT) {
            cf acc{0.0f, 0.0f};
            const float2 *xp = xs + pos;
            if (CPLX) {  // ComplexFirFilter.java:147-151: acc += (tr + j ti) * (xr + j xi)
#pragma unroll 4
                for (int t = 0; t < a.ntaps; t++) {
                    const float2 h = st[t], x = xp[-t];
                    acc = caxpy(h.x, cf{x.x, x.y}, acc);
                    acc = caxpy(h.y, cf{-x.y, x.x}, acc);
                }
            } else if (a.real_only) {  // FirFilter.kt:141-146
#pragma unroll 4
                for (int t = 0; t < a.ntaps; t++) acc.x = fmaf(st[t].x, xp[-t].x, acc.x);
            } else {  // FirFilter.kt:90-96
#pragma unroll 4
                for (int t = 0; t < a.ntaps; t++) {
                    const float2 x = xp[-t];
                    acc = caxpy(st[t].x, cf{x.x, x.y}, acc);
                }
            }
            ar = acc.x;
            ai = acc.y;
        } else if (CPLX) {
            for (int t = 0; t < a.ntaps; t++) {
                const float tr = str_[t], ti = sti[t], xr = sre[pos - t], xi = sim[pos - t];
                ar = __fadd_rn(ar, __fsub_rn(__fmul_rn(tr, xr), __fmul_rn(ti, xi)));
                ai = __fadd_rn(ai, __fadd_rn(__fmul_rn(ti, xr), __fmul_rn(tr, xi)));
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
    pdl_enter();
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
                            long long nout, float *out_re, float *out_im, bool exact, cudaStream_t st, int rs_span,
                            float *hist_new_re, float *hist_new_im, long long consumed, bool *hist_done) {
    if (hist_done) *hist_done = false;
    if (nout <= 0) return cudaSuccess;
    if (nt + 2 > kSpanMax) return cudaErrorInvalidValue;
    ResampleArgs a{};
    const bool want_hist = hist_done && hist_new_re && hist_new_im && in.hist > 0 && consumed > 0;
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
    const bool aligned = in.kind == 3 || ((size_t)in.raw & (in.kind == 2 ? 3 : 1)) == 0;  // one IQ pair per load
    const int A = (nt + D - 1) / D;
    // stripe path (large decimation: many taps per output window, windows of a phase far apart)
    if (!exact && aligned && nt >= 256 && 2 * D >= nt && I <= 16) {
        constexpr int kSpanS = 11300;  // samples per CTA (90 KB) + their raw codes (45 KB): one persistent CTA per SM
        constexpr int PH = 2, TP = 8;  // two phases x eight periods per warp task; twelve warps per CTA share the tasks' sample ranges
        long long tpc = ((long long)kSpanS - nt - D - 24) / D;
        tpc -= tpc % TP;
        if (tpc >= TP) {
            ResampleStripeArgs sa{};
            sa.a = a;
            if (want_hist) {
                sa.a.hist_new_re = hist_new_re;
                sa.a.hist_new_im = hist_new_im;
                sa.a.consumed = consumed;
            }
            sa.TPC = (int)tpc;
            sa.pad = (int)(((long long)(PH - 1) * D + I - 1) / I) + 1 + 68;
            sa.a.tile = (int)(tpc * I);
            sa.a.span_max = (int)(tpc * D + nt + D + 24);
            const int groups = (I + PH - 1) / PH, tasks = groups * (int)(tpc / TP);
            (void)tasks;
            // default: two persistent CTAs of twelve warps per SM, raw IQ loaded by the threads (knob rs_span = 1: eight warps);
            // variant B (knob rs_span = 2): one CTA of twelve warps per SM, the next tile's raw IQ arrives by bulk copy
            const bool variant_b = rs_span == 2;
            const bool variant_d = rs_span != 1 && rs_span != 2;  // the default; knob rs_span = 1: eight warps per CTA (97 registers)
            const int warps = (variant_b || variant_d) ? 12 : 8;
            sa.part_pairs = (int)((tpc * I + 1) & ~1LL);
            sa.acc_pairs = sa.part_pairs * (tasks < warps ? warps / tasks : 1);  // the kernel's nsplit
            sa.raw_bytes = (variant_b && in.kind <= 2) ? (int)(((size_t)sa.a.span_max * (in.kind == 2 ? 4 : 2) + 15) & ~(size_t)15) : 0;
            sa.nco_pairs = in.nco_cos ? ((in.nco_len + 1 + 7) & ~7) : 8;  // one entry more than the table: its entry 0 again
            const size_t ssmem = ((size_t)sa.a.span_max + 10 + sa.nco_pairs + sa.acc_pairs) * sizeof(float2) + (size_t)I * (nt + 2 * sa.pad) * sizeof(float) +
                                 (size_t)(2 * I + 2) * sizeof(int) + 16 + sa.raw_bytes;
            if (ssmem <= (variant_b ? 200 : 113) * 1024) {
                long long stiles = (nout + sa.a.tile - 1) / sa.a.tile;
                int sms = 148;
                {
                    int dev = 0;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                }
                const long long resident = variant_b ? sms : 2LL * sms;
                const unsigned sgrid = (unsigned)(stiles < resident ? stiles : resident);  // persistent CTAs walk the tiles
                static DeviceOnce sonce;
                int sdev = 0;
                const bool sfirst = sonce.pending(&sdev);
                const int mx = 200 * 1024;
                const bool dodd = (D & 1) != 0;
#define RFA_STRIPE(KIND, DO)                                                                                                    \
    do {                                                                                                                        \
        if (sfirst) cudaFuncSetAttribute(resample_stripe_kernel<KIND, PH, TP, DO>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx); \
        if (sfirst) cudaFuncSetAttribute(resample_stripe12_kernel<KIND, PH, TP, DO>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx); \
        if (launch_now && in.kind == KIND && dodd == DO) {                                                                      \
            if (variant_d)                                                                                                      \
                resample_stripe12_kernel<KIND, PH, TP, DO><<<sgrid, 32 * warps, ssmem, st>>>(sa);                               \
            else                                                                                                                \
                resample_stripe_kernel<KIND, PH, TP, DO><<<sgrid, 32 * warps, ssmem, st>>>(sa);                                 \
        }                                                                                                                       \
    } while (0)
                if (in.kind < 0 || in.kind > 3) return cudaErrorInvalidValue;
                for (int pass = sfirst ? 0 : 1; pass < 2; pass++) {
                    const bool launch_now = pass == 1;
                    RFA_STRIPE(0, false); RFA_STRIPE(0, true);
                    RFA_STRIPE(1, false); RFA_STRIPE(1, true);
                    RFA_STRIPE(2, false); RFA_STRIPE(2, true);
                    RFA_STRIPE(3, false); RFA_STRIPE(3, true);
                    if (pass == 0) sonce.done(sdev);
                }
#undef RFA_STRIPE
                if (want_hist) *hist_done = true;
                return cudaGetLastError();
            }
        }
    }
    if (!exact && aligned && A >= 2 && A <= 12 && (size_t)I * D * 12 <= 8192) {
        // register-tiled path: M = 9 outputs per thread, B blocks per phase
        // 4 K samples (+ bank + NCO table = 40 KB) per CTA: B = 16 blocks per phase, so that the 32 tasks of a warp share ONE
        // phase and their tap loads are broadcasts; registers (64 x 256) allow four CTAs per SM, shared memory five
        constexpr int M = 9, kSpanT = 4096;
        const int span_cap = rs_span > 0 && rs_span < kSpanT ? rs_span : kSpanT;  // tuning knob "rs_span"
        const int amax = A <= 3 ? 3 : (A <= 5 ? 5 : (A <= 9 ? 9 : 12));
        long long B = ((long long)span_cap - 8 - (long long)(amax + 1) * D) / ((long long)M * D);  // span <= (B*M + amax + 1)*D + 4
        if (B >= 16) B -= B % 16;
        if (B >= 1) {
            ResampleTiledArgs ta{};
            ta.a = a;
            if (want_hist) {
                ta.a.hist_new_re = hist_new_re;
                ta.a.hist_new_im = hist_new_im;
                ta.a.consumed = consumed;
            }
            ta.a.span_max = kSpanT;
            ta.a.tile = (int)((long long)I * M * B);
            ta.B = (int)B;
            ta.A = A;
            ta.AP = (amax + 3) & ~3;
            const unsigned tgrid = (unsigned)((nout + ta.a.tile - 1) / ta.a.tile);
            long long threads = ((long long)I * B + 31) / 32 * 32;
            if (threads > 256) threads = 256;
            const int span_used = (int)((B * M + amax + 1) * D + 8);
            ta.a.span_max = span_used;
            ta.nco_pairs = (in.nco_cos && in.kind <= 2) ? ((in.nco_len + 1 + 7) & ~7) : 0;
            const size_t tsmem = ((size_t)span_used + 8) * sizeof(float2) + (size_t)I * D * ta.AP * sizeof(float) +
                                 (size_t)ta.nco_pairs * sizeof(float2);
            static DeviceOnce tonce;
            int tdev = 0;
            const bool tfirst = tonce.pending(&tdev);
            const int mx = (int)(((size_t)kSpanT + 8) * sizeof(float2) + 8192 * sizeof(float) + 512 * sizeof(float2));
#define RFA_RT(KIND, AM)                                                                                           \
    do {                                                                                                           \
        if (tfirst)                                                                                                 \
            cudaFuncSetAttribute(resample_tiled_kernel<KIND, M, AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx); \
        if (launch_now) resample_tiled_kernel<KIND, M, AM><<<tgrid, (unsigned)threads, tsmem, st>>>(ta);            \
    } while (0)
#define RFA_RT_ALL(AM)                               \
    do {                                             \
        RFA_RT(0, AM);                               \
        RFA_RT(1, AM);                               \
        RFA_RT(2, AM);                               \
        RFA_RT(3, AM);                               \
    } while (0)
            if (tfirst) {  // opt every instantiation in to the large shared-memory carve-out once per device
                const bool launch_now = false;
                RFA_RT_ALL(3);
                RFA_RT_ALL(5);
                RFA_RT_ALL(9);
                RFA_RT_ALL(12);
                tonce.done(tdev);
            }
            {
                const bool launch_now = true;
#define RFA_RT_KIND(AM)                                  \
    switch (in.kind) {                                   \
        case 0: RFA_RT(0, AM); break;                    \
        case 1: RFA_RT(1, AM); break;                    \
        case 2: RFA_RT(2, AM); break;                    \
        case 3: RFA_RT(3, AM); break;                    \
        default: return cudaErrorInvalidValue;           \
    }
                if (amax == 3) { RFA_RT_KIND(3) }
                else if (amax == 5) { RFA_RT_KIND(5) }
                else if (amax == 9) { RFA_RT_KIND(9) }
                else { RFA_RT_KIND(12) }
#undef RFA_RT_KIND
            }
#undef RFA_RT_ALL
#undef RFA_RT
            if (want_hist) *hist_done = true;
            return cudaGetLastError();
        }
    }
    if (!exact && aligned) {
        ResampleFastArgs fa{};
        fa.a = a;
        // lanes per output: one when a tile has enough outputs for every thread, otherwise 16 -- a
        // half-warp per output reads 16 neighbouring samples, which is conflict-free whatever the
        // distance between the windows of neighbouring outputs
        const int G = (tile >= 256 && nout >= 256) ? 1 : 16;
        fa.G = G;
        // the bank goes to shared memory only when copying it is cheap next to the tile's work
        const size_t bank_floats = (size_t)I * nt;
        fa.bank_smem = bank_floats * sizeof(float) <= 4 * 1024 ? (int)bank_floats : 0;
        const size_t fsmem = ((size_t)kSpanMax + 8) * sizeof(float2) + (size_t)fa.bank_smem * sizeof(float);
        static DeviceOnce fonce;
        int fdev = 0;
        const bool ffirst = fonce.pending(&fdev);
        const int mx = (int)(((size_t)kSpanMax + 8) * sizeof(float2) + 4 * 1024);
#define RFA_RF1(KIND, BS, GG)                                                                                          \
    do {                                                                                                               \
        if (ffirst)                                                                                                    \
            cudaFuncSetAttribute(resample_fast_kernel<KIND, BS, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx); \
        if (launch_now && in.kind == KIND && (fa.bank_smem != 0) == BS && G == GG)                                     \
            resample_fast_kernel<KIND, BS, GG><<<grid, 256, fsmem, st>>>(fa);                                          \
    } while (0)
#define RFA_RF(KIND)                 \
    do {                             \
        RFA_RF1(KIND, true, 1);      \
        RFA_RF1(KIND, true, 16);     \
        RFA_RF1(KIND, false, 1);     \
        RFA_RF1(KIND, false, 16);    \
    } while (0)
        if (in.kind < 0 || in.kind > 3) return cudaErrorInvalidValue;
        for (int pass = ffirst ? 0 : 1; pass < 2; pass++) {
            const bool launch_now = pass == 1;
            RFA_RF(0);
            RFA_RF(1);
            RFA_RF(2);
            RFA_RF(3);
            if (pass == 0) fonce.done(fdev);
        }
#undef RFA_RF1
#undef RFA_RF
        return cudaGetLastError();
    }
#define RFA_RS(KIND)                                                                              \
    do {                                                                                          \
        if (exact)                                                                                \
            resample_kernel<KIND, true><<<grid, 256, smem, st>>>(a);                              \
        else                                                                                      \
            resample_kernel<KIND, false><<<grid, 256, smem, st>>>(a);                             \
    } while (0)
    static DeviceOnce ronce;
    int rdev = 0;
    if (ronce.pending(&rdev)) {
        cudaFuncSetAttribute(resample_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resample_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ronce.done(rdev);
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
    // a short stream (the quadrature-rate filters behind a large decimation: 10^5 outputs per call) in 1024-output
    // tiles is 79 CTAs on 148 SMs -- four CTAs per SM at least, down to one output per thread
    {
        static thread_local int sm_dev = -1, sms = 148;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev != sm_dev) {
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
            sm_dev = dev;
        }
        long long fill = (nout + 4LL * sms - 1) / (4LL * sms);
        fill = (fill + 255) / 256 * 256;
        if (fill < tile) tile = fill;
    }
    a.tile = (int)tile;
    a.real_only = real_only ? 1 : 0;
    a.out_re = out_re;
    a.out_im = out_im;
    const unsigned grid = (unsigned)((nout + tile - 1) / tile);
    const size_t smem = (2 * (size_t)kSpanMax + 2 * (size_t)ntaps) * sizeof(float);
    static DeviceOnce konce;
    int kdev = 0;
    if (konce.pending(&kdev)) {
        const int mx = (int)((2 * (size_t)kSpanMax + 2 * 4096) * sizeof(float));
        cudaFuncSetAttribute(fir_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(fir_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        konce.done(kdev);
    }
    if (taps_im) {
        if (exact)
            pdl_launch(fir_kernel<true, true>, grid, 256, smem, st, a);
        else
            pdl_launch(fir_kernel<true, false>, grid, 256, smem, st, a);
    } else {
        if (exact)
            pdl_launch(fir_kernel<false, true>, grid, 256, smem, st, a);
        else
            pdl_launch(fir_kernel<false, false>, grid, 256, smem, st, a);
    }
    return cudaGetLastError();
}

cudaError_t history_launch(const StreamDesc &in, long long consumed, float *new_re, float *new_im,
                           cudaStream_t st) {
    if (in.hist <= 0) return cudaSuccess;
    const StreamSrc s = make_src(in);
    const unsigned grid = (unsigned)((in.hist + 127) / 128);
    switch (in.kind) {
        case 0: pdl_launch(history_kernel<0>, grid, 128, 0, st, s, consumed, new_re, new_im); break;
        case 1: pdl_launch(history_kernel<1>, grid, 128, 0, st, s, consumed, new_re, new_im); break;
        case 2: pdl_launch(history_kernel<2>, grid, 128, 0, st, s, consumed, new_re, new_im); break;
        case 3: pdl_launch(history_kernel<3>, grid, 128, 0, st, s, consumed, new_re, new_im); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace rfa

#ifdef RFA_STRIPE_TRACE
extern "C" int rfa_debug_stripe_trace(unsigned long long *out) {
    cudaDeviceSynchronize();
    unsigned long long z[8] = {0};
    cudaMemcpyFromSymbol(out, rfa::g_stripe_trace, sizeof(z));
    cudaMemcpyToSymbol(rfa::g_stripe_trace, z, sizeof(z));
    return 0;
}
#endif
