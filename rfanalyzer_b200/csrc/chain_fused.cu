// chain_fused.cu -- everything behind the resampler of an FM chain in ONE kernel:
//   Demodulator.applyUserFilter   (A/analyzer/Demodulator.kt:215-240)  27-tap low pass on the quadrature samples
//   Demodulator.demodulateFM      (:251-275)                           gain * atan2 of x[n] * conj(x[n-1]), one sample of carry
//   Demodulator.run volume        (:184-187)
//   AudioSink.applyAudioFilter    (A/analyzer/AudioSink.java:215-237)  the /2 (and /4) audio decimators
// Round 1 ran them as four FIR / demodulator launches plus a state kernel each: nine launches
// behind the resampler for a few MB of data.  Here a CTA owns a run of TILE demodulator samples: it stages the
// quadrature samples that run needs (with the halo of every filter behind it), runs the user filter, the
// discriminator and the decimators out of shared memory, and writes each intermediate sample it OWNS once -- the
// demodulated stream and the first decimator's output still go to global memory, because they are the delay lines
// ("history") of the next call.  The arithmetic per output is the same as fir.cu / demod.cu: taps in the reference's
// order, SUM_EXACT with separately rounded products and sums and a double-precision atan2.
//
// A second small kernel slides every delay line of the chain in one launch.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"
#include "rfa_fft_core.cuh"  // cf, caxpy (packed FFMA2)
#include "pdl.h"

namespace rfa {
namespace {

// atan2 for the RFA_SUM_FMA discriminator: odd minimax polynomial of degree 17 on [0, 1] (|error| < 1.1e-7 rad in
// float32, fitted in tools/fit_atan.py), one approximate division, octant folding -- a quarter of atan2f's instructions,
// which were a third of this kernel.  RFA_SUM_EXACT keeps the double-precision atan2 of Math.atan2.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    const float t2 = t * t;
    float p = 0.0024567203962780693f;
    p = fmaf(p, t2, -0.014401341682745893f);
    p = fmaf(p, t2, 0.039781197940081274f);
    p = fmaf(p, t2, -0.07234855251230274f);
    p = fmaf(p, t2, 0.10498945099958955f);
    p = fmaf(p, t2, -0.14161228944659016f);
    p = fmaf(p, t2, 0.19985906733990874f);
    p = fmaf(p, t2, -0.3333259702641045f);
    p = fmaf(p, t2, 0.9999998863828075f);
    float r = p * t;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.0f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float a, float b) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(a, b)) : fmaf(a, b, acc);
}

// TILE = demodulator samples owned by a CTA: 2048, or 512 when the call is too short to fill the machine with those
constexpr int HALO = 40;       // >= 1 (discriminator) + 8 (first decimator) + 2 * 12 (second) demodulator samples before the run
constexpr int UT = 32;         // user filter taps at most (27 in the reference)

template <bool EXACT, int TILE>
__global__ void __launch_bounds__(256) fm_tail_kernel(const FmTailArgs a) {
    pdl_enter();
    // quadrature samples (user filter input) as (re, im) pairs; eight entries of slack in front: the last tap chunk of a
    // thread's first outputs looks (never uses) a few entries before the first staged sample
    __shared__ float2 s_q[8 + TILE + HALO + UT + 16];
    __shared__ float s_ure[TILE + HALO], s_uim[TILE + HALO];            // user filter output
    __shared__ float s_dem[TILE + HALO];                                // demodulated (index 0 = u0 - HALO + 1)
    __shared__ float s_a1[(TILE + HALO) / 2 + 2];                       // first decimator output
    __shared__ __align__(16) float s_tu[UT];
    __shared__ float s_t1[16], s_t2[16];
    if (a.slide_user && blockIdx.x == gridDim.x - 1) {  // the extra CTA: the user filter's next delay line
        const ChainStateArgs::Line &l = a.user_line;
        for (int h = threadIdx.x; h < l.hist; h += blockDim.x) {
            const long long k = l.n - l.hist + h;
            float r = 0.0f, q = 0.0f;
            if (k >= 0) {
                r = l.in_re[k];
                q = l.in_im[k];
            } else if (k + l.hist >= 0) {
                r = l.old_re[k + l.hist];
                q = l.old_im[k + l.hist];
            }
            l.new_re[h] = r;
            l.new_im[h] = q;
        }
        return;
    }
    const long long u0 = (long long)blockIdx.x * TILE;                  // first demodulator index this CTA owns
    long long u1 = u0 + TILE;
    if (u1 > a.nu) u1 = a.nu;
    if (u0 >= u1) return;
    if (blockIdx.x == 0) {  // calls shorter than a decimator's delay line keep the newest of the old one
        if (a.a1_hist_new)
            for (int h = threadIdx.x; h < a.a1_hist - a.nu; h += blockDim.x) a.a1_hist_new[h] = a.hist_a1[a.nu + h];
        if (a.a2_hist_new)
            for (int h = threadIdx.x; h < a.a2_hist - a.n1; h += blockDim.x) a.a2_hist_new[h] = a.hist_a2[a.n1 + h];
    }
    for (int t = threadIdx.x; t < UT; t += blockDim.x) s_tu[t] = t < a.user_taps ? a.taps_user[t] : 0.0f;
    if (threadIdx.x < a.a1_taps) s_t1[threadIdx.x] = a.taps_a1[threadIdx.x];
    if (threadIdx.x < a.a2_taps) s_t2[threadIdx.x] = a.taps_a2[threadIdx.x];
    if (threadIdx.x < 8) s_q[threadIdx.x] = make_float2(0.0f, 0.0f);
    // ---- quadrature samples: user output i needs inputs first_u + i - (taps-1) .. first_u + i (decimation 1) ----
    const long long ulo = u0 - HALO;                                    // first user output computed here (may be < 0)
    const long long qlo = a.first_u + ulo - (a.user_taps - 1);
    const int nqs = (int)(u1 - ulo) + a.user_taps - 1;
    float2 *xq = s_q + 8;
    auto stage_one = [&](int s) {
        const long long k = qlo + s;
        float r = 0.0f, q = 0.0f;
        if (k >= 0) {
            r = a.q_re[k];
            q = a.q_im[k];
        } else if (k + a.user_hist >= 0) {
            r = a.hist_u_re[k + a.user_hist];
            q = a.hist_u_im[k + a.user_hist];
        }
        xq[s] = make_float2(r, q);
    };
    if (qlo >= 0) {  // every CTA but the call's first: eight loads in flight per thread instead of a chain of round trips
        const float *gr = a.q_re + qlo, *gi = a.q_im + qlo;
        int s = threadIdx.x;
        for (; s + 3 * (int)blockDim.x < nqs; s += 4 * blockDim.x) {
            float r[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                r[u] = __ldg(gr + s + u * blockDim.x);
                q[u] = __ldg(gi + s + u * blockDim.x);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) xq[s + u * blockDim.x] = make_float2(r[u], q[u]);
        }
        for (; s < nqs; s += blockDim.x) stage_one(s);
    } else {
        for (int s = threadIdx.x; s < nqs; s += blockDim.x) stage_one(s);
    }
    __syncthreads();
    // ---- user filter (FirFilter.kt:90-96), outputs ulo .. u1-1; outputs with index < 0 belong to earlier calls ----
    // A thread owns M consecutive outputs and walks the taps eight at a time, in the reference's order (t = 0 first):
    // output s, tap t reads sample s + taps-1 - t, so a chunk of eight taps needs a window of M + 7 samples -- one
    // 64-bit load per (re, im) pair and M * 8 multiply-adds per chunk instead of three loads per multiply-add pair.
    const int nus = (int)(u1 - ulo);
    constexpr int M = (TILE + HALO + 255) / 256;  // 9 (TILE 2048) / 3 (TILE 512): odd, so that lanes M samples apart hit distinct banks
    static_assert(M % 2 == 1, "M must be odd");
    for (int s0 = threadIdx.x * M; s0 < nus; s0 += blockDim.x * M) {
        float ar[M], ai[M];  // (RFA_SUM_FMA keeps the pair in `ac` instead: one packed multiply-add per tap and output)
        cf ac[M];
#pragma unroll
        for (int m = 0; m < M; m++) {
            ar[m] = ai[m] = 0.0f;
            ac[m] = cf{0.0f, 0.0f};
        }
        for (int t0 = 0; t0 < a.user_taps; t0 += 8) {
            const float2 *xw = xq + (s0 + a.user_taps - 8 - t0);  // sample of (output s0, tap t0 + 7)
            float2 w[M + 7];
#pragma unroll
            for (int i = 0; i < M + 7; i++) w[i] = xw[i];
            const float4 ha = *reinterpret_cast<const float4 *>(s_tu + t0), hb = *reinterpret_cast<const float4 *>(s_tu + t0 + 4);
            const float h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (t0 + j < a.user_taps) {  // CTA-uniform
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        if constexpr (EXACT) {
                            ar[m] = mac<EXACT>(ar[m], h[j], w[7 + m - j].x);
                            ai[m] = mac<EXACT>(ai[m], h[j], w[7 + m - j].y);
                        } else {
                            ac[m] = caxpy(h[j], cf{w[7 + m - j].x, w[7 + m - j].y}, ac[m]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; m++) {
            const int sm = s0 + m;
            if (sm < nus) {
                float vr = EXACT ? ar[m] : ac[m].x, vi = EXACT ? ai[m] : ac[m].y;
                if (ulo + sm == -1) {  // the sample before this call's first one: the discriminator's carry
                    vr = a.carry_in[0];
                    vi = a.carry_in[1];
                } else if (ulo + sm < 0) {
                    vr = vi = 0.0f;
                }
                s_ure[sm] = vr;
                s_uim[sm] = vi;
            }
        }
    }
    __syncthreads();
    // ---- discriminator (Demodulator.kt:262-270) for demodulator indices ulo+1 .. u1-1 ----
    for (int s = threadIdx.x + 1; s < nus; s += blockDim.x) {
        const long long i = ulo + s;
        float d = 0.0f;
        if (i >= 0) {
            const float r = s_ure[s], q = s_uim[s], pr = s_ure[s - 1], pq = s_uim[s - 1];
            if (EXACT) {
                const float x = __fadd_rn(__fmul_rn(r, pr), __fmul_rn(q, pq));
                const float y = __fsub_rn(__fmul_rn(q, pr), __fmul_rn(r, pq));
                d = __fmul_rn(__fmul_rn(a.gain, (float)atan2((double)y, (double)x)), a.volume);
            } else {
                d = a.gain * fast_atan2(fmaf(q, pr, -(r * pq)), fmaf(r, pr, q * pq)) * a.volume;
            }
            if (i >= u0) {
                if (a.dem_out) a.dem_out[i] = d;
                if (a.ratio == 1) a.audio[i] = d;
                if (a.a1_hist_new && i >= a.nu - a.a1_hist) a.a1_hist_new[i - (a.nu - a.a1_hist)] = d;
            }
        } else if (i + a.a1_hist >= 0 && a.hist_a1) {
            d = a.hist_a1[i + a.a1_hist];  // demodulated samples of earlier calls: the first decimator's delay line
        }
        s_dem[s] = d;
    }
    if (threadIdx.x == 0 && u1 == a.nu) {  // Demodulator.kt:271-272: the newest filtered sample is the next call's carry
        a.carry_out[0] = s_ure[nus - 1];
        a.carry_out[1] = s_uim[nus - 1];
    }
    if (a.ratio == 1) return;
    __syncthreads();
    // ---- first decimator (FirFilter.kt:141-146): output m <-> demodulator index first_a1 + 2*m ----
    // computed here: every m whose newest tap lies in (ulo + 8, u1); owned: newest tap in [u0, u1)
    const long long dlo = ulo + 1;  // demodulator index of s_dem[1]
    long long m_lo = (dlo + (a.a1_taps - 1) - a.first_a1 + 1) / 2;  // ceil((dlo + taps-1 - first) / 2)
    if ((dlo + (a.a1_taps - 1) - a.first_a1) < 0) m_lo = -((-(dlo + (a.a1_taps - 1) - a.first_a1)) / 2);
    long long m_hi = (u1 - 1 - a.first_a1) >= 0 ? (u1 - 1 - a.first_a1) / 2 : -1;  // last m with newest tap <= u1-1
    const int nm = (int)(m_hi - m_lo + 1);
    // A thread owns MD = 5 consecutive outputs (their windows slide by two samples: 17 loads for 45 multiply-adds instead of
    // two loads per multiply-add; an odd MD keeps the lanes' 10-float stride at a two-way bank conflict, like the stride of
    // two it replaces), taps in registers, in the reference's order.
    constexpr int MD = 5;
    float h1[9];
#pragma unroll
    for (int t = 0; t < 9; t++) h1[t] = t < a.a1_taps ? s_t1[t] : 0.0f;
    for (int s0 = threadIdx.x * MD; s0 < nm; s0 += blockDim.x * MD) {
        const int pos0 = (int)(a.first_a1 + 2 * (m_lo + s0) - dlo) + 1;  // s_dem index of output s0's newest sample
        float w[2 * MD + 7];
#pragma unroll
        for (int i = 0; i < 2 * MD + 7; i++) {
            const int idx = pos0 - 8 + i;
            w[i] = (idx >= 0 && idx < TILE + HALO) ? s_dem[idx] : 0.0f;
        }
        float acc[MD];
#pragma unroll
        for (int k = 0; k < MD; k++) acc[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; t++) {
            if (t < a.a1_taps) {  // CTA-uniform
#pragma unroll
                for (int k = 0; k < MD; k++) acc[k] = mac<EXACT>(acc[k], h1[t], w[8 + 2 * k - t]);
            }
        }
#pragma unroll
        for (int k = 0; k < MD; k++) {
            const int sk = s0 + k;
            if (sk < nm) {
                const long long m = m_lo + sk;
                float v = 0.0f;
                if (m >= 0) {
                    v = acc[k];
                    if (a.first_a1 + 2 * m >= u0 && m < a.n1) {
                        if (a.a1_out) a.a1_out[m] = v;
                        if (a.ratio == 2) a.audio[m] = v;
                        if (a.a2_hist_new && m >= a.n1 - a.a2_hist) a.a2_hist_new[m - (a.n1 - a.a2_hist)] = v;
                    }
                } else if (m + a.a2_hist >= 0 && a.hist_a2) {
                    v = a.hist_a2[m + a.a2_hist];  // first-decimator outputs of earlier calls: the second decimator's delay line
                }
                s_a1[sk] = v;
            }
        }
    }
    if (a.ratio != 8) return;
    __syncthreads();
    // ---- second decimator: output b <-> first-decimator index first_a2 + 4*b; owned when that sample's newest
    // demodulator index lies in [u0, u1) ----
    if (nm <= 0) return;
    // b range whose window [newest-12, newest] lies inside [m_lo, m_hi] and whose newest sample is owned
    long long b_lo = (m_lo + (a.a2_taps - 1) - a.first_a2 + 3) / 4;
    if ((m_lo + (a.a2_taps - 1) - a.first_a2) < 0) b_lo = -((-(m_lo + (a.a2_taps - 1) - a.first_a2)) / 4);
    if (b_lo < 0) b_lo = 0;
    const long long b_hi = (m_hi - a.first_a2) >= 0 ? (m_hi - a.first_a2) / 4 : -1;
    for (long long b = b_lo + threadIdx.x; b <= b_hi && b < a.n2; b += blockDim.x) {
        const long long newest_m = a.first_a2 + 4 * b;
        const long long newest_d = a.first_a1 + 2 * newest_m;
        if (newest_d < u0) continue;  // owned by the CTA before this one
        const int pos = (int)(newest_m - m_lo);
        float acc = 0.0f;
        for (int t = 0; t < a.a2_taps; t++) acc = mac<EXACT>(acc, s_t2[t], s_a1[pos - t]);
        a.audio[b] = acc;
    }
}

// Slides the delay lines of the user filter and of both decimators in one launch (FirFilter keeps the newest
// taps-1 inputs): new = last `hist` samples of (old history ++ this call's `n` inputs).
__global__ void chain_state_kernel(const ChainStateArgs a) {
    pdl_enter();
    const int which = blockIdx.x;  // 0: user filter (complex), 1: first decimator, 2: second decimator
    const ChainStateArgs::Line &l = a.line[which];
    for (int h = threadIdx.x; h < l.hist; h += blockDim.x) {
        const long long k = l.n - l.hist + h;  // stream index relative to this call's first input
        float r = 0.0f, q = 0.0f;
        if (k >= 0) {
            r = l.in_re[k];
            if (l.in_im) q = l.in_im[k];
        } else if (k + l.hist >= 0) {
            r = l.old_re[k + l.hist];
            if (l.old_im) q = l.old_im[k + l.hist];
        }
        l.new_re[h] = r;
        if (l.new_im) l.new_im[h] = q;
    }
}

}  // namespace

cudaError_t fm_tail_launch(const FmTailArgs &a, bool exact, cudaStream_t st) {
    if (a.nu <= 0) return cudaSuccess;
    if (a.user_taps > UT || a.user_taps < 1 || a.a1_taps > 9 || a.a2_taps > 13) return cudaErrorInvalidValue;
    const unsigned extra = a.slide_user ? 1u : 0u;
    if (a.nu >= 2048LL * 512) {
        const unsigned grid = (unsigned)((a.nu + 2047) / 2048) + extra;
        if (exact)
            pdl_launch(fm_tail_kernel<true, 2048>, grid, 256, 0, st, a);
        else
            pdl_launch(fm_tail_kernel<false, 2048>, grid, 256, 0, st, a);
    } else {
        const unsigned grid = (unsigned)((a.nu + 511) / 512) + extra;
        if (exact)
            pdl_launch(fm_tail_kernel<true, 512>, grid, 256, 0, st, a);
        else
            pdl_launch(fm_tail_kernel<false, 512>, grid, 256, 0, st, a);
    }
    return cudaGetLastError();
}

cudaError_t chain_state_launch(const ChainStateArgs &a, cudaStream_t st) {
    pdl_launch(chain_state_kernel, 3, 64, 0, st, a);
    return cudaGetLastError();
}

}  // namespace rfa
