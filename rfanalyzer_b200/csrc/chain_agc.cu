// chain_agc.cu -- everything behind the resampler of an AM / LSB / USB / CW chain (RFA_SUM_FMA) in two or three kernels:
//   agc_tail_kernel   Demodulator.applyUserFilter (A/analyzer/Demodulator.kt:215-240), then
//                       AM:        re*re + im*im                              (:293-297)
//                       SSB / CW:  complex band-pass, real part               (:325-345, :372-385, ComplexFirFilter.java:147-151)
//                     and the per-packet maximum (and sum, AM) the AGC needs  (:290-296, :347-349)
//   agc_scan_kernel   lastMax = max(0.95 * lastMax, packet maximum) over the packets of the call (:299-302, :350-354);
//                     up to 1024 packets per call the apply kernel's CTAs run the recurrence themselves
//   agc_apply_kernel  (x - mean) * gain * volume, for AM the /2 audio decimator (AudioSink.java:215-237), and the
//                     call's bookkeeping (delay lines, table of the next call)
// Round 1 ran this as nine to ten launches (two FIR kernels with a history kernel each, power, packet statistics, scan,
// normalise, audio decimator + history, a device copy) behind a host-synchronised segment upload: 55 us per 2^24-sample
// call for 1.6e5 quadrature samples.  (The packet boundaries now come from packet_table_kernel, in closed form.)  Here a CTA of the first kernel owns a run of TILE demodulated samples: it stages the
// quadrature samples that run needs, runs the user filter into shared memory (de-interleaved by the band filter's
// decimation phase), the band-pass out of shared memory, and adds its packets' maxima to the call's table with one atomic
// per packet and warp.  The 181-tap complex band-pass was bound by shared-memory wavefronts (a tap and a sample load per
// multiply-add pair); here a thread owns MB consecutive outputs and walks the taps eight at a time: MB + 7 sample loads and
// four 128-bit tap loads per 8 * MB packed multiply-adds, and only the real part is formed (the reference computes the
// imaginary part too and drops it, Demodulator.kt:343).
// RFA_SUM_EXACT chains keep the separate kernels (fir.cu, demod.cu): their sums follow the reference's order.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"
#include "pdl.h"
#include "rfa_fft_core.cuh"

namespace rfa {
namespace {

constexpr int UT = 32;          // user filter taps at most (27 in the reference)
constexpr int MB = 3;           // consecutive band-pass outputs per thread (odd: lanes MB samples apart hit distinct banks)
constexpr int TILE = 256 * MB;  // demodulated samples per CTA
constexpr int MU = 7;           // consecutive user-filter outputs per thread (odd; a tile's outputs are one round of the CTA)
constexpr int A1T = 9;          // first audio decimator: taps at most

// elementwise packed multiply-add (one FFMA2): (a.x * b.x + c.x, a.y * b.y + c.y)
__device__ __forceinline__ cf fma_elem(cf a, cf b, cf c) {
#ifdef RFA_PACKED
    return fma2(RFA_PK(a), RFA_PK(b), RFA_PK(c));
#else
    return cf{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
}

// float -> unsigned key with the same order (0 = below everything, the table's cleared state)
__device__ __forceinline__ unsigned enc_max(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_max(unsigned k) {
    if (k == 0u) return -INFINITY;  // an empty packet: the old kernels' fmaxf over nothing
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// largest p with off[p] <= i (off[0] = 0, off[npk] > i); empty packets are stepped over.  Packets are (nearly) equal in
// length, so the proportional guess is right or one off almost always: its three neighbouring boundaries are loaded at once
// and only a miss pays for the binary search's chain of dependent loads.
__device__ __forceinline__ int find_packet(const long long *__restrict__ off, int npk, long long i) {
    const long long total = __ldg(off + npk);
    int g = total > 0 ? (int)((double)i * (double)npk / (double)total) : 0;
    g = g < 1 ? 1 : (g > npk - 1 ? npk - 1 : g);
    if (npk >= 2) {
        const long long o0 = __ldg(off + g - 1), o1 = __ldg(off + g), o2 = __ldg(off + g + 1);
        if (o1 <= i && i < o2) return g;
        if (o0 <= i && i < o1 && o0 < o1) return g - 1;
    }
    int lo = 0, hi = npk;  // off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// shared-memory layout of agc_tail_kernel (float2 units), computed the same way by the launcher
struct TailLayout {
    int kp;      // taps per decimation phase, rounded up to a multiple of 8
    int ulen;    // entries of one de-interleaved user-output array
    int nq_max;  // staged quadrature samples at most
    int total;   // float2 entries in all
};
__host__ __device__ inline TailLayout tail_layout(int band_taps, int dec, int user_taps) {
    TailLayout l;
    const int per_phase = band_taps > 0 ? (band_taps + dec - 1) / dec : 0;
    l.kp = (per_phase + 7) & ~7;
    // outputs j < TILE look at U_r[j + k], k < kp; the call's last CTA carries up to dec - 1 samples past its last window
    l.ulen = band_taps > 0 ? TILE + l.kp + 8 : TILE + 8;
    const int nus_max = band_taps > 0 ? dec * (TILE - 1) + band_taps + dec : TILE;
    l.nq_max = nus_max + user_taps - 1;
    l.total = 8 + l.nq_max + MU + 16 + (band_taps > 0 ? dec * l.ulen + dec * l.kp : l.ulen);
    return l;
}

template <int DEC, bool BAND>
__global__ void __launch_bounds__(256) agc_tail_kernel(const AgcTailArgs a) {
    pdl_enter();
    extern __shared__ float2 smem_tail[];
    __shared__ float s_x[TILE];
    __shared__ __align__(16) float s_tu[UT];
    const TailLayout L = tail_layout(BAND ? a.band_taps : 0, DEC, a.user_taps);
    float2 *xq = smem_tail + 8;                       // quadrature samples (eight entries of slack in front)
    float2 *s_U = smem_tail + 8 + L.nq_max + MU + 16; // [DEC][ulen] user-filter outputs by decimation phase
    float2 *s_G = s_U + DEC * L.ulen;                 // [DEC][kp]   band-pass taps by phase, (re, -im), reversed
    const long long x0 = (long long)blockIdx.x * TILE;
    long long x1 = x0 + TILE;
    if (x1 > a.nx) x1 = a.nx;
    if (x0 >= x1) return;
    const bool last_cta = x1 == a.nx;
    __shared__ int s_plo;  // packet of the CTA's first sample: looked up now, its loads overlap the staging
    if (threadIdx.x == 255) s_plo = find_packet(a.off, a.npk, x0);
    for (int t = threadIdx.x; t < UT; t += blockDim.x) s_tu[t] = t < a.user_taps ? a.taps_user[t] : 0.0f;
    if (threadIdx.x < 8) smem_tail[threadIdx.x] = make_float2(0.0f, 0.0f);
    if (BAND) {
        // Out_j = sum_t h[t] * u[first_b + DEC*j - t] = sum_r sum_k G_r[k] * U_r[j + k],  e = taps-1-t = DEC*k + r
        for (int i = threadIdx.x; i < DEC * L.kp; i += blockDim.x) {
            const int r = i / L.kp, k = i - r * L.kp;
            const int t = a.band_taps - 1 - (DEC * k + r);
            s_G[i] = t >= 0 ? make_float2(a.taps_b_re[t], -a.taps_b_im[t]) : make_float2(0.0f, 0.0f);
        }
        for (int i = threadIdx.x; i < DEC * L.ulen; i += blockDim.x) s_U[i] = make_float2(0.0f, 0.0f);
    }
    // ---- user-filter outputs this CTA needs: global indices u_lo .. u_hi ----
    long long u_lo, u_hi;
    if (BAND) {
        u_lo = a.first_b + DEC * x0 - (a.band_taps - 1);
        u_hi = last_cta ? a.nu - 1 : a.first_b + DEC * (x1 - 1);  // the samples behind the call's last window are the next call's delay line
    } else {
        u_lo = x0;
        u_hi = x1 - 1;
    }
    const int nus = (int)(u_hi - u_lo + 1);
    const long long qlo = a.first_u + u_lo - (a.user_taps - 1);
    const int nqs = nus + a.user_taps - 1;
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
    if (qlo >= 0) {
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
    for (int s = nqs + threadIdx.x; s < nqs + MU + 16; s += blockDim.x) xq[s] = make_float2(0.0f, 0.0f);
    __syncthreads();
    // ---- user filter (FirFilter.kt:90-96): a thread owns MU consecutive outputs, taps eight at a time ----
    for (int s0 = threadIdx.x * MU; s0 < nus; s0 += blockDim.x * MU) {
        cf ac[MU];
#pragma unroll
        for (int m = 0; m < MU; m++) ac[m] = cf{0.0f, 0.0f};
        for (int t0 = 0; t0 < a.user_taps; t0 += 8) {
            const float2 *xw = xq + (s0 + a.user_taps - 8 - t0);  // sample of (output s0, tap t0 + 7)
            float2 w[MU + 7];
#pragma unroll
            for (int i = 0; i < MU + 7; i++) w[i] = xw[i];
            const float4 ha = *reinterpret_cast<const float4 *>(s_tu + t0), hb = *reinterpret_cast<const float4 *>(s_tu + t0 + 4);
            const float h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int j = 0; j < 8; j++)  // taps past user_taps are zero
#pragma unroll
                for (int m = 0; m < MU; m++) ac[m] = caxpy(h[j], cf{w[7 + m - j].x, w[7 + m - j].y}, ac[m]);
        }
#pragma unroll
        for (int m = 0; m < MU; m++) {
            const int sm = s0 + m;
            if (sm < nus) {
                float vr = ac[m].x, vi = ac[m].y;
                const long long gi = u_lo + sm;
                if (BAND) {
                    if (gi < 0) {  // outputs of earlier calls: the band-pass's delay line
                        vr = a.hist_b_re[gi + a.band_hist];
                        vi = a.hist_b_im[gi + a.band_hist];
                    } else if (gi >= a.nu - a.band_hist) {  // the next call's delay line (neighbouring CTAs write equal values)
                        a.u_out_re[gi] = vr;
                        a.u_out_im[gi] = vi;
                    }
                    s_U[(sm % DEC) * L.ulen + sm / DEC] = make_float2(vr, vi);
                } else {
                    s_x[sm] = fmaf(vr, vr, vi * vi);  // Demodulator.kt:293-297
                }
            }
        }
    }
    __syncthreads();
    const int nxs = (int)(x1 - x0);
    if (BAND) {
        // ---- complex band-pass, real part: a thread owns outputs MB*tid .. MB*tid + MB-1 ----
        const int j0 = threadIdx.x * MB;
        if (j0 < nxs) {
            cf acc[MB];
#pragma unroll
            for (int m = 0; m < MB; m++) acc[m] = cf{0.0f, 0.0f};
#pragma unroll
            for (int r = 0; r < DEC; r++) {
                const float2 *U = s_U + r * L.ulen + j0;
                const float4 *G = reinterpret_cast<const float4 *>(s_G + r * L.kp);
                for (int k0 = 0; k0 < L.kp; k0 += 8) {
                    float2 w[MB + 7];
#pragma unroll
                    for (int i = 0; i < MB + 7; i++) w[i] = U[k0 + i];
                    float4 g4[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) g4[i] = G[(k0 >> 1) + i];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const cf g = (q & 1) ? cf{g4[q >> 1].z, g4[q >> 1].w} : cf{g4[q >> 1].x, g4[q >> 1].y};
#pragma unroll
                        for (int m = 0; m < MB; m++) acc[m] = fma_elem(g, cf{w[m + q].x, w[m + q].y}, acc[m]);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < MB; m++)
                if (j0 + m < nxs) s_x[j0 + m] = acc[m].x + acc[m].y;
        }
        __syncthreads();
    }
    // ---- demodulated samples out, per-packet maximum (and sum) ----
    for (int j = threadIdx.x; j < nxs; j += blockDim.x) a.x_out[x0 + j] = s_x[j];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p_lo = s_plo;
    for (int p = p_lo + warp; p < a.npk; p += 8) {
        const long long b = __ldg(a.off + p), e = __ldg(a.off + p + 1);
        if (b >= x1) break;
        const int lo = (int)((b > x0 ? b : x0) - x0), hi = (int)((e < x1 ? e : x1) - x0);
        if (lo >= hi) continue;
        float m = -INFINITY, s = 0.0f;
        for (int j = lo + lane; j < hi; j += 32) {
            const float v = s_x[j];
            m = fmaxf(m, v);
            s += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
            s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        }
        if (lane == 0) {
            atomicMax(a.mx_enc + p, enc_max(m));
            if (!BAND) atomicAdd(a.sum + p, (double)s);
        }
    }
}

// Demodulator.kt:290,296,299-302 (AM) and :347-355 (SSB/CW): the AGC recurrence over packets, as demod.cu's scan kernel,
// reading the table agc_tail_kernel filled.  Only calls of more than SCAN_INLINE packets launch it: below that every CTA
// of agc_apply_kernel runs the recurrence itself (a few cycles per packet) and a launch is saved.
constexpr int SCAN_INLINE = 1024;
__global__ void __launch_bounds__(256) agc_scan_enc_kernel(const long long *off, int npackets, const double *sum,
                                                           const unsigned *mx_enc, const float *state_in, float *state_out,
                                                           float *gain, float *mean) {
    pdl_enter();
    __shared__ float s_last[1024];
    __shared__ float s_carry;
    if (threadIdx.x == 0) s_carry = state_in[0];
    for (int p0 = 0; p0 < npackets; p0 += 1024) {
        const int np = npackets - p0 < 1024 ? npackets - p0 : 1024;
        for (int i = threadIdx.x; i < np; i += blockDim.x) s_last[i] = dec_max(mx_enc[p0 + i]);
        __syncthreads();
        if (threadIdx.x == 0) {
            float last = s_carry;
            for (int i = 0; i < np; i++) {
                last = __fmul_rn(last, (float)0.95);
                const float m = s_last[i];
                if (m > last) last = m;
                s_last[i] = last;
            }
            s_carry = last;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const int p = p0 + i;
            gain[p] = __fdiv_rn(0.75f, s_last[i]);
            const long long n = off[p + 1] - off[p];
            mean[p] = __fdiv_rn((float)sum[p], (float)n);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) state_out[0] = s_carry;
}

// new delay line = the last `hist` samples of (old delay line ++ this call's n inputs); chain_fused.cu's state kernel
__device__ __forceinline__ void slide_line(const ChainStateArgs::Line &l) {
    for (int h = threadIdx.x; h < l.hist; h += blockDim.x) {
        const long long k = l.n - l.hist + h;
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

// out = (x - mean[p]) * gain[p] * volume (AM) or x * gain[p] * volume (SSB / CW); RATIO 2: the normalised samples are
// the input of the first audio decimator (FirFilter.kt:141-146), whose outputs are the call's audio.  The kernel also
// finishes the call's bookkeeping: two extra CTAs slide the delay lines of the user filter and the band-pass, the CTAs
// that own the newest normalised samples write the decimator's next delay line, and CTA 0 clears the packet table
// the NEXT call will fill (the two tables alternate; this call's table is still being read by the other CTAs).
constexpr int CT = 1024;  // demodulated samples per CTA
template <bool SUBTRACT_MEAN, int RATIO>
__global__ void __launch_bounds__(256) agc_apply_kernel(const AgcApplyArgs a) {
    pdl_enter();
    __shared__ float s_y[CT + A1T];
    __shared__ float s_t1[A1T];
    __shared__ float s_gain[SCAN_INLINE], s_mean[SCAN_INLINE];
    const unsigned ntiles = (unsigned)((a.nx + CT - 1) / CT);
    if (blockIdx.x >= ntiles) {
        slide_line(a.line[blockIdx.x - ntiles]);
        return;
    }
    const long long c0 = (long long)blockIdx.x * CT;
    long long c1 = c0 + CT;
    if (c1 > a.nx) c1 = a.nx;
    // this thread's demodulated samples: loaded now, so that their latency hides behind the packet search and the scan
    constexpr int XPT = (CT + A1T + 255) / 256;
    const long long x_lo = c0 - (RATIO == 2 ? a.a1_taps - 1 : 0);  // RATIO 2: s_y[0]
    float xv[XPT];
#pragma unroll
    for (int k = 0; k < XPT; k++) {
        const long long i = x_lo + threadIdx.x + 256 * k;
        xv[k] = (i >= 0 && i < c1) ? __ldg(a.x + i) : 0.0f;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < a.clear_n; i += blockDim.x) {
            a.clear_mx[i] = 0u;
            a.clear_sum[i] = 0.0;
        }
        if (RATIO == 2)  // a call shorter than the decimator's delay line keeps the newest of the old one
            for (int h = threadIdx.x; h < a.a1_hist - a.nx; h += blockDim.x) a.a1_hist_new[h] = a.hist_a1[a.nx + h];
    }
    const bool last_cta = blockIdx.x == ntiles - 1;
    const float *gain = a.gain, *mean = a.mean;
    __shared__ int s_prange[2];  // packets of this CTA's first (halo) and last sample
    if (threadIdx.x < 2) {
        long long i = threadIdx.x == 0 ? c0 - (RATIO == 2 ? A1T : 0) : c1 - 1;
        s_prange[threadIdx.x] = find_packet(a.off, a.npk, i < 0 ? 0 : i);
    }
    __syncthreads();
    const int p_first = s_prange[0], p_last = s_prange[1];
    if (a.scan_inline) {
        // packets 0 .. p_end-1: what this CTA's samples need (the last CTA runs to the end and leaves the AGC state)
        const int p_end = last_cta ? a.npk : p_last + 1;
        for (int p = threadIdx.x; p < p_end; p += blockDim.x) {
            s_gain[p] = dec_max(a.mx_enc[p]);
            if (SUBTRACT_MEAN && p >= p_first) s_mean[p] = __fdiv_rn((float)a.sum[p], (float)(a.off[p + 1] - a.off[p]));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // last = max(0.95 * last, m): the only sequential part, eight packets' maxima in registers at a time (the
            // maximum is never NaN -- fmaxf drops NaN samples -- so fmaxf equals the reference's `if (m > last)`)
            float last = a.state_in[0];
            for (int p0 = 0; p0 < p_end; p0 += 8) {
                float m[8];
#pragma unroll
                for (int u = 0; u < 8; u++) m[u] = p0 + u < p_end ? s_gain[p0 + u] : 0.0f;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (p0 + u < p_end) last = fmaxf(__fmul_rn(last, (float)0.95), m[u]);
                    m[u] = last;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (p0 + u < p_end) s_gain[p0 + u] = m[u];
            }
            if (last_cta) a.state_out[0] = last;
        }
        __syncthreads();
        for (int p = p_first + threadIdx.x; p <= p_last; p += blockDim.x) s_gain[p] = __fdiv_rn(0.75f, s_gain[p]);
        __syncthreads();
        gain = s_gain;
        mean = s_mean;
    }
    // the packet of a sample: a binary search between the CTA's first and last packet (mostly one or two steps)
    auto packet_of = [&](long long i) {
        int lo = p_first, hi = p_last + 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.off + mid) <= i)
                lo = mid;
            else
                hi = mid;
        }
        return lo;
    };
    auto normalised = [&](long long i, float v) {
        const int p = packet_of(i);
        if (SUBTRACT_MEAN) v = __fsub_rn(v, mean[p]);
        return __fmul_rn(__fmul_rn(v, gain[p]), a.volume);
    };
    if (RATIO == 1) {
#pragma unroll
        for (int k = 0; k < XPT; k++) {
            const long long i = x_lo + threadIdx.x + 256 * k;
            if (i < c1) a.audio[i] = normalised(i, xv[k]);
        }
        return;
    }
    if (threadIdx.x < A1T) s_t1[threadIdx.x] = threadIdx.x < a.a1_taps ? a.taps_a1[threadIdx.x] : 0.0f;
    const long long lo = x_lo;  // s_y[0]
#pragma unroll
    for (int k = 0; k < XPT; k++) {
        const int s = threadIdx.x + 256 * k;
        const long long i = lo + s;
        if (i >= c1) continue;
        float y = 0.0f;
        if (i >= 0) {
            y = normalised(i, xv[k]);
            if (i >= c0 && i >= a.nx - a.a1_hist) a.a1_hist_new[i - (a.nx - a.a1_hist)] = y;  // the decimator's next delay line
        } else if (i + a.a1_hist >= 0) {
            y = a.hist_a1[i + a.a1_hist];
        }
        s_y[s] = y;
    }
    __syncthreads();
    // decimator outputs whose newest sample first_a1 + 2*m lies in [c0, c1)
    long long m_lo = c0 - a.first_a1 <= 0 ? 0 : (c0 - a.first_a1 + 1) / 2;
    for (long long m = m_lo + threadIdx.x; m < a.n1; m += blockDim.x) {
        const long long newest = a.first_a1 + 2 * m;
        if (newest >= c1) break;
        const int pos = (int)(newest - lo);
        float acc = 0.0f;
        for (int t = 0; t < a.a1_taps; t++) acc = fmaf(s_t1[t], s_y[pos - t], acc);
        a.audio[m] = acc;
    }
}

// the call's packet boundaries, one thread per packet (launched BEFORE the resampler: the three kernels behind it then
// follow each other without a copy in between)
__global__ void packet_table_kernel(const PacketMap pm, long long *off) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= pm.npk) off[p] = pm.off(p);
}

}  // namespace

cudaError_t packet_table_launch(const PacketMap &pm, long long *off, cudaStream_t st) {
    if (pm.npk < 0) return cudaErrorInvalidValue;
    packet_table_kernel<<<(unsigned)((pm.npk + 1 + 255) / 256), 256, 0, st>>>(pm, off);
    return cudaGetLastError();
}

bool agc_tail_supported(const AgcTailArgs &a) {
    if (a.nx <= 0 || a.user_taps < 1 || a.user_taps > UT || a.npk < 1) return false;
    if (a.band_taps > 0) {
        if (a.band_dec != 1 && a.band_dec != 2) return false;
        if (a.band_taps > 1024 || a.band_hist != a.band_taps - 1) return false;
    }
    return true;
}

cudaError_t agc_tail_launch(const AgcTailArgs &a, cudaStream_t st) {
    if (!agc_tail_supported(a)) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((a.nx + TILE - 1) / TILE);
    const TailLayout L = tail_layout(a.band_taps, a.band_taps > 0 ? a.band_dec : 1, a.user_taps);
    const size_t smem = (size_t)L.total * sizeof(float2);
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kernel) {
        if (smem > 40 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = pdl_launch(kernel, grid, 256, smem, st, a);
    };
    if (a.band_taps <= 0)
        go(agc_tail_kernel<1, false>);
    else if (a.band_dec == 1)
        go(agc_tail_kernel<1, true>);
    else
        go(agc_tail_kernel<2, true>);
    return e != cudaSuccess ? e : cudaGetLastError();
}

bool agc_scan_is_inline(int npk) { return npk <= SCAN_INLINE; }

cudaError_t agc_scan_enc_launch(const long long *off, int npk, const double *sum, const unsigned *mx_enc, const float *state_in,
                                float *state_out, float *gain, float *mean, cudaStream_t st) {
    if (npk <= 0) return cudaSuccess;
    pdl_launch(agc_scan_enc_kernel, 1, 256, 0, st, off, npk, sum, mx_enc, state_in, state_out, gain, mean);
    return cudaGetLastError();
}

cudaError_t agc_apply_launch(const AgcApplyArgs &a, bool subtract_mean, cudaStream_t st) {
    if (a.nx <= 0) return cudaErrorInvalidValue;
    if ((a.ratio != 1 && a.ratio != 2) || a.a1_taps > A1T || a.nlines < 0 || a.nlines > 2) return cudaErrorInvalidValue;
    if (a.scan_inline && a.npk > SCAN_INLINE) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((a.nx + CT - 1) / CT) + (unsigned)a.nlines;
    if (a.ratio == 1) {
        if (subtract_mean)
            pdl_launch(agc_apply_kernel<true, 1>, grid, 256, 0, st, a);
        else
            pdl_launch(agc_apply_kernel<false, 1>, grid, 256, 0, st, a);
    } else {
        if (subtract_mean)
            pdl_launch(agc_apply_kernel<true, 2>, grid, 256, 0, st, a);
        else
            pdl_launch(agc_apply_kernel<false, 2>, grid, 256, 0, st, a);
    }
    return cudaGetLastError();
}

}  // namespace rfa
