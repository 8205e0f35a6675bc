// spectrum_kernel.cuh -- the fused IQ -> spectrum kernel (K3 of SURVEY.md 2b).
//
// One launch does, per FFT frame, what the reference spreads over three threads:
//   Scheduler.kt:266          fillPacketIntoSamplePacket   bytes -> float re/im   (LUT)
//   NativeDsp.kt:55-58        re*w, im*w                   Blackman window
//   nativedsp.cpp:69          pffft_transform_ordered      forward C2C FFT
//   nativedsp.cpp:72-79       10*log10(sqrt((re/N)^2+(im/N)^2)) at (i+N/2)%N
//   FftProcessor.kt:224       waterfall row store
//   FftProcessor.kt:244       peaks = max(peaks, row)
// so an IQ sample crosses HBM once (2 or 4 bytes in) and one float per bin goes out.
//
// Layout: a CTA slot of T threads owns one frame of NL points in a padded shared
// frame; each thread keeps E = NL/T points in registers.  CTAs are persistent: the grid
// is sized to the machine and every slot walks frames slot, slot+G, slot+2G, ...  The
// bins a thread produces in the last pass do not depend on the frame, so the peak-hold
// running maximum lives in registers for the whole launch and is written once.
//
// N > 16384 does not fit one SM's shared memory.  It is split N = S * NL (S = 2, 4):
// CTA residue c computes bins k = c + S*k2 by folding the S strided input sections with
// W_S^(n1*c), rotating by W_N^(n2*c) and running the NL-point transform -- the raw input
// (2-4 B/sample) is re-read S times from L2, nothing is exchanged between CTAs.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

#include "rfa_fft_core.cuh"

namespace rfa {

enum : int { OUT_DB = 0, OUT_CPLX = 1 };

struct SpectrumParams {
    const void *in;       // raw IQ (s8/u8 pairs, s16le pairs, interleaved cf32) or planar re
    const float *in_im;   // planar im (FMT_PF32 only)
    const float *win;     // [N] window, nullptr = rectangular
    const cf *tw;         // per-pass Stockham twiddles (make_pass_twiddles(NL))
    const cf *twN;        // [N]  exp(-2*pi*i*t/N), only read when S > 1
    float *rows;          // OUT_DB: dB rows; OUT_CPLX: interleaved complex spectra
    long long row0;       // output row of frame f = (row0 + f*row_step) mod ring_rows
    long long row_step;
    long long ring_rows;  // 0 = no wrap
    long long row_stride; // floats between rows
    long long nframes;
    long long store_from; // rows of frames < store_from are not written (peak-only runs)
    float *peaks;         // [N] running maxima, merged with float atomics at kernel end, or nullptr
    float *avg;           // [N] mean of the newest avg_len+1 rows (computed by the tail CTA group), or nullptr
    long long avg_newest; // row index of the newest row after this launch
    long long avg_dir;    // step towards older rows (-row_step)
    long long avg_valid;  // rows that hold data; older terms count as -9999f
    int avg_len;          // L
    unsigned int *ticket; // [8] zeroed counters: [c] finished tail rows, [4+c] CTAs done averaging (residue c < 4)
    float inv_n2;         // dB bias -3.0103*log2(N) (the 1/N^2 scaling, applied after the logarithm)
    int pdl;              // launched with programmatic stream serialization: constant tables first, then wait for the previous grid
};

template <int NL>
struct Geom {
#ifdef RFA_R8
    static constexpr int T = NL == 4096 ? 512 : (NL >= 8192 ? 512 : (NL >= 16 * 16 ? NL / 16 : (NL / 16 > 0 ? NL / 16 : 1)));
#elif !defined(RFA_8192_T512)
    static constexpr int T = NL == 8192 ? 256 : (NL >= 8192 ? 512 : (NL >= 16 * 16 ? NL / 16 : (NL / 16 > 0 ? NL / 16 : 1)));
#else
    static constexpr int T = NL >= 8192 ? 512 : (NL >= 16 * 16 ? NL / 16 : (NL / 16 > 0 ? NL / 16 : 1));
#endif
    static constexpr int E = NL / T;
    static constexpr int FPC = T >= 256 ? 1 : 256 / T;   // frames per CTA
    static constexpr int CTA = T * FPC;
    // two ping-pong frames per slot (one barrier per pass) while they fit comfortably
#ifndef RFA_8192_T512
    static constexpr int NBUF = NL < 8192 ? 2 : 1;
#else
    static constexpr int NBUF = NL <= 8192 ? 2 : 1;
#endif
    static constexpr size_t SMEM = (size_t)FPC * NBUF * Plan<NL>::SMEM_POINTS * sizeof(cf);
};

RFA_HD cf rot_q(cf v, int q) {  // v * (-j)^q
    q &= 3;
    if (q == 1) return cf{v.y, -v.x};
    if (q == 2) return cf{-v.x, -v.y};
    if (q == 3) return cf{-v.y, v.x};
    return v;
}

// one input point (index n inside frame f of N points), converted and multiplied by
// ws = window tap * unit scale of the format (see unit_scale)
template <int IN>
RFA_CX float unit_scale() {
    return IN == FMT_S16LE ? (1.0f / 32768.0f) : ((IN == FMT_S8 || IN == FMT_U8) ? 0.0078125f : 1.0f);
}

// integer code pair in a register -> (I, Q) * ws
template <int IN>
RFA_HD cf decode_point(uint32_t raw, float ws) {
    cf m;
    if (IN == FMT_S8) {
        raw ^= 0x8080u;  // two's complement -> offset binary: code + 128
        m = cf{magic_byte0(raw), magic_byte1(raw)};
        return cscale(cadd(m, cf{-8388736.0f, -8388736.0f}), ws);
    } else if (IN == FMT_U8) {
        m = cf{magic_byte0(raw), magic_byte1(raw)};
        return cscale(cadd(cadd(m, cf{-8388608.0f, -8388608.0f}), cf{-127.4f, -127.4f}), ws);
    } else {
        raw ^= 0x80008000u;
        m = cf{magic_half0(raw), magic_half1(raw)};
        return cscale(cadd(m, cf{-8421376.0f, -8421376.0f}), ws);
    }
}

template <int IN>
RFA_HD cf load_point(const void *in, const float *in_im, long long idx, float ws) {
    cf v;
    if (IN == FMT_S8 || IN == FMT_U8) {
        v = decode_point<IN>((uint32_t)((const uint16_t *)in)[idx], ws);
    } else if (IN == FMT_S16LE) {
        v = decode_point<IN>(((const uint32_t *)in)[idx], ws);
    } else if (IN == FMT_CF32) {
        v = cscale(((const cf *)in)[idx], ws);
    } else {
        v.x = ((const float *)in)[idx] * ws;
        v.y = in_im[idx] * ws;
    }
    return v;
}

template <int IN>
RFA_CX int in_elem_bytes() {
    return IN == FMT_S16LE ? 4 : (IN == FMT_CF32 ? 8 : (IN == FMT_PF32 ? 4 : 2));
}

// Everything one thread does for one frame between two barriers is a "phase"; the
// kernel and the CPU emulation (tests/emu) call the same phase functions.
template <int NL, int S, int IN, int OUT>
struct SpectrumFrame {
    using G = Geom<NL>;
    using PL = Plan<NL>;
    static constexpr int T = G::T, E = G::E, N = NL * S;
    static constexpr int LAST = PL::PASSES - 1;
    // raw codes of the next frame are fetched while the current one is transformed
    static constexpr bool PREFETCH = (S == 1) && (E <= 16) && (IN == FMT_S8 || IN == FMT_U8 || IN == FMT_S16LE);
    // twiddles of the last pass live in registers when there are few enough of them
    static constexpr int LAST_TW = PL::PASSES > 1 ? (E / PL::radix(LAST)) * (PL::radix(LAST) - 1) : 0;
    static constexpr bool LAST_TW_REG = LAST_TW > 0 && LAST_TW <= 16;
    // entries of a last-pass butterfly's register twiddles that are kept (composed_twiddle)
    // measured per size (profiles/r02k_twiddles_kept_sweep.txt, us per 2^24 samples, all / six / two kept): N = 256: 34.1 /
    // 35.7 / 31.4, 2048: 44.1 / 41.8 / 41.1, 4096: 43.8 / 40.5 / 42.9, 128: no difference
#ifdef RFA_TWKEEP
    static constexpr int LAST_TW_KEEP = RFA_TWKEEP;
#else
    static constexpr int LAST_TW_KEEP = NL == 4096 ? 6 : 2;
#endif
    // radix-4 last pass (N = 64, 1024): W^2k, W^3k from the one kept W^k (1024: 41.3 -> 39.7 us, same profile file)
#ifdef RFA_R4KEEP_ALL
    static constexpr bool RADIX4_KEEP1 = false;
#else
    static constexpr bool RADIX4_KEEP1 = true;
#endif
    // twiddle tables of the middle passes are staged in shared memory (<= 32 KB)
    static constexpr int MID_TW = PL::PASSES > 2 ? pass_tw_offset<NL>(LAST) : 0;
    static constexpr bool MID_TW_SMEM = MID_TW > 0 && MID_TW <= 4096;
    static constexpr size_t SMEM_BYTES = G::SMEM + (MID_TW_SMEM ? MID_TW * sizeof(cf) : 0);

    // phase 0a: fetch the raw codes of frame f (integer formats)
    // `src` points at point `tid` of the frame
    static RFA_HD void load_raw(const char *src, uint32_t *raw) {
        constexpr int R = PL::radix(0);
        constexpr int NB = E / R, STR = NL / R;
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int off = b * T + r * STR;
#ifdef RFA_EXP_NOLOAD
                raw[b * R + r] = (uint32_t)(size_t)src + off;
                continue;
#endif
                raw[b * R + r] = (IN == FMT_S16LE) ? ((const uint32_t *)src)[off] : (uint32_t)((const uint16_t *)src)[off];
            }
    }
    // phase 0b: convert + window, first radix pass in registers
    static RFA_HD void first_from_raw(const uint32_t *raw, const float *wreg, cf *u) {
        constexpr int R = PL::radix(0);
#pragma unroll
        for (int e = 0; e < E; e++) u[e] = decode_point<IN>(raw[e], wreg[e]);
        pass_first_compute<NL, T, R>(u);
    }

    // phase 0 (non-prefetching formats and split transforms): gather + convert + window + fold
    template <bool SH = false>
    static RFA_HD void first(const SpectrumParams &p, long long f, int c, int tid, const float *wreg, cf *u,
                             const char *staged = nullptr) {
        constexpr int R = PL::radix(0);
        constexpr int NB = E / R, STR = NL / R;
        // one 64-bit address per frame; every point of this thread is a constant offset from it
        // (`staged`: the frame's raw codes in shared memory instead -- STAGED kernels with 32 points per thread)
        // SH: the frame's raw codes sit in shared memory (`staged`); the compiler is told so, or the loads are generic LDs
        const char *src = (SH ? staged : (const char *)p.in + f * (long long)N * in_elem_bytes<IN>()) + (size_t)tid * in_elem_bytes<IN>();
#ifdef __CUDA_ARCH__
        if constexpr (SH) __builtin_assume(__isShared(src));
#endif
        const float *src_im = (IN == FMT_PF32) ? p.in_im + (f * (long long)N + tid) : nullptr;
#pragma unroll
        for (int b = 0; b < NB; b++) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int off = b * T + r * STR;
                cf acc;
                if (S == 1) {
                    acc = load_point<IN>(src, src_im, off, wreg[b * R + r]);
                } else {
                    const int n2 = tid + off;
                    acc = cf{0.f, 0.f};
#pragma unroll
                    for (int n1 = 0; n1 < S; n1++) {
                        const float ws = p.win ? p.win[n2 + n1 * NL] * unit_scale<IN>() : unit_scale<IN>();
                        const cf v = load_point<IN>(src, src_im, off + n1 * NL, ws);
                        acc = cadd(acc, rot_q(v, n1 * c * (4 / S)));
                    }
                    if (c != 0) acc = cmul(acc, p.twN[(n2 * c) & (N - 1)]);
                }
                u[b * R + r] = acc;
            }
        }
        pass_first_compute<NL, T, R>(u);
    }

    template <int PASS>
    static RFA_HD void scatter(cf *x, int tid, const cf *u) {
        pass_scatter<NL, T, PL::radix(PASS), PL::prod(PASS)>(x, tid, u);
    }
    // `tw` is the base of the per-pass tables (global or the shared-memory copy)
    template <int PASS>
    static RFA_HD void gather(const cf *x, const cf *tw, int tid, cf *u) {
#ifndef RFA_NO_TWCOMPOSE
        // both radix-32 passes of the 16 x 32 x 32 plan (N = 16384): 64.8 -> 60.1 us (int8), 74.1 -> 69.6 us (int16) per 2^24
        // samples, profiles/r02h_16384_composed_twiddles.txt
        if constexpr (PASS > 0 && PL::radix(PASS) == 32 && E == 32 && !(PASS == LAST && LAST_TW_REG))
            pass_gather_r32_composed<NL, T, PL::prod(PASS)>(x, tw + pass_tw_offset<NL>(PASS), tid, u);
        else
#endif
        // middle radix-16 passes: 6 twiddle loads and 9 products instead of 15 loads (N = 8192: 58.8 -> 57.0 us, 4096: 44.3 ->
        // 43.9 us, 2048: 45.0 -> 44.4 us per 2^24 samples, profiles/r02k_composed_radix16_twiddles.txt)
        if constexpr (PASS > 0 && PASS != LAST && PL::radix(PASS) == 16 && PL::prod(PASS) >= 16 && (NL / PL::radix(PASS)) % 16 == 0)
            pass_gather_r16_composed<NL, T, PL::prod(PASS)>(x, tw + pass_tw_offset<NL>(PASS), tid, u);
        else
            pass_gather<NL, T, PL::radix(PASS), PL::prod(PASS)>(x, tw + pass_tw_offset<NL>(PASS), tid, u);
    }
    // last pass with its twiddles held by the thread: twreg[b*(R-1) + r-1]
    static RFA_HD void load_last_tw(const cf *tw, int tid, cf *twreg) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST), NB = E / R;
        const cf *t = tw + pass_tw_offset<NL>(LAST);
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int r = 1; r < R; r++) twreg[b * (R - 1) + r - 1] = t[(r - 1) * P + ((tid + b * T) & (P - 1))];
    }
    // a middle pass with its twiddles fetched into registers by the caller (before the exchange barrier)
    template <int PASS>
    static RFA_HD void load_pass_tw(const cf *tw, int tid, cf *twreg) {
        constexpr int R = PL::radix(PASS), P = PL::prod(PASS), NB = E / R;
        const cf *t = tw + pass_tw_offset<NL>(PASS);
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int r = 1; r < R; r++) twreg[b * (R - 1) + r - 1] = t[(r - 1) * P + ((tid + b * T) & (P - 1))];
    }
    template <int PASS>
    static RFA_HD void gather_reg(const cf *x, const cf *twreg, int tid, cf *u) {
        constexpr int R = PL::radix(PASS), NB = E / R, STR = NL / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const cf *xi = x + phys(tid + b * T);
#pragma unroll
            for (int r = 0; r < R; r++) {
                cf v = xi[r * (STR + STR / 16)];
                if (r > 0) v = cmul(v, twreg[b * (R - 1) + r - 1]);
                u[b * R + r] = v;
            }
            Dft<R>::run(u + b * R);
        }
    }
    static RFA_HD void gather_last_reg(const cf *x, const cf *twreg, int tid, cf *u) {
        constexpr int R = PL::radix(LAST), NB = E / R, STR = NL / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const cf *xi = x + phys(tid + b * T);
#pragma unroll
            for (int r = 0; r < R; r++) {
#ifdef RFA_EXP_NOXCHG
                cf v = u[b * R + r];
#else
                cf v = xi[r * (STR + STR / 16)];
#endif
                if (r > 0) {
                    if constexpr (R == 16 || R == 8) {
                        v = cmul(v, composed_twiddle<R, LAST_TW_KEEP>(twreg + b * (R - 1), r));
                    } else if constexpr (R == 4 && RADIX4_KEEP1) {  // W^2k, W^3k from W^k: one kept entry per butterfly
                        const cf w1 = twreg[b * 3], w2 = cmul(w1, w1);
                        v = cmul(v, r == 1 ? w1 : (r == 2 ? w2 : cmul(w2, w1)));
                    } else {
                        v = cmul(v, twreg[b * (R - 1) + r - 1]);
                    }
                }
                u[b * R + r] = v;
            }
            Dft<R>::run(u + b * R);
        }
    }

    // final phase: dB (or raw complex) to the output row, running peak maximum.
    // `out` is the row base (row index resolved by the caller).  Bin k = c + S*(tid + b*T + cc*P)
    // and tid + b*T < P <= N/2/S, so the fft-shift XOR only touches the constant part.
    template <bool PEAK, bool STORE>
    static RFA_HD void emit(float *out, int c, int tid, const cf *u, float *pk, float inv_n2) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST), NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (OUT == OUT_CPLX) {
                cf *o = (cf *)out + (c + S * (tid + b * T));
#pragma unroll
                for (int cc = 0; cc < R; cc++)
                    if (STORE) o[S * cc * P] = u[b * R + Dft<R>::perm(cc)];
            } else {
                float *o = out + (c + S * (tid + b * T));
#ifdef RFA_DB2
                if (R % 2 == 0) {
#pragma unroll
                    for (int cc = 0; cc < R; cc += 2) {
                        float d0, d1;
                        logmag_db2(u[b * R + Dft<R>::perm(cc)], u[b * R + Dft<R>::perm(cc + 1)], inv_n2, d0, d1);
                        if (STORE) o[shifted_offset(S * cc * P)] = d0;
                        if (STORE) o[shifted_offset(S * (cc + 1) * P)] = d1;
                        if (PEAK) pk[b * R + cc] = fmaxf(pk[b * R + cc], d0);
                        if (PEAK) pk[b * R + cc + 1] = fmaxf(pk[b * R + cc + 1], d1);
                    }
                    continue;
                }
#endif
#pragma unroll
                for (int cc = 0; cc < R; cc++) {
#ifdef RFA_EXP_NOEMIT
                    pk[b * R + cc] = fmaxf(pk[b * R + cc], u[b * R + Dft<R>::perm(cc)].x);
                    continue;
#endif
                    const float db = logmag_db(u[b * R + Dft<R>::perm(cc)], inv_n2);
                    if (STORE) o[shifted_offset(S * cc * P)] = db;
                    if (PEAK) pk[b * R + cc] = fmaxf(pk[b * R + cc], db);
                }
            }
        }
    }
    // offset of bin (base + d) after the fft-shift, relative to base, for d a multiple of
    // S*P with base < S*P: (base + d) ^ (N/2) = base + (d ^ (N/2)) when N/2 is a multiple of S*P
    static RFA_CX int shifted_offset(int d) {
        return (PL::prod(PL::PASSES - 1) * S <= N / 2) ? (d ^ (N >> 1)) : d;
    }
    // shifted output index of peak register e (= b*R + cc) of thread tid
    static RFA_HD int peak_index(int c, int tid, int e) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST);
        const int b = e / R, cc = e % R;
        return (c + S * (tid + b * T + cc * P)) ^ (N >> 1);
    }
};

// row of frame f: (row0 + f*row_step) mod ring_rows
RFA_HD long long frame_row(const SpectrumParams &p, long long f) {
    long long row = p.row0 + f * p.row_step;
    if (p.ring_rows > 0) {
        row %= p.ring_rows;
        if (row < 0) row += p.ring_rows;
    }
    return row;
}

// AnalyzerSurface.kt:683-684,710-714 for bin i: newest -> oldest, float32, divided by L+1
template <class LoadFn>
RFA_HD float boxcar_average(const SpectrumParams &p, int i, LoadFn load) {
    float sum = 0.0f;
    for (int r = 0; r <= p.avg_len; r++) {
        float v = -9999.0f;
        if (r < p.avg_valid) {
            long long row = p.avg_newest + (long long)r * p.avg_dir;
            if (p.ring_rows > 0) {
                row %= p.ring_rows;
                if (row < 0) row += p.ring_rows;
            }
            v = load(p.rows + row * p.row_stride + i);
        }
        sum = sum + v;
    }
    return sum / (float)(p.avg_len + 1);
}

#ifdef __CUDACC__
// RFA_TRACE (tuning builds only): thread 0 of every CTA stamps clock64() at phase boundaries into
// g_trace[cta][slot]; tools read it back through rfa_debug_trace().
#ifdef RFA_TRACE
#define RFA_TRACE_SLOTS 128
__device__ long long g_trace[512][RFA_TRACE_SLOTS];
#define RFA_STAMP(k)                                                                          \
    do {                                                                                      \
        if (threadIdx.x == 0 && (k) < RFA_TRACE_SLOTS) g_trace[blockIdx.x & 511][(k)] = clock64(); \
    } while (0)
#else
#define RFA_STAMP(k) \
    do {             \
    } while (0)
#endif

// max for floats of either sign through the integer atomics (RED.MAX / RED.MIN)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.0f)
        atomicMax((int *)addr, __float_as_int(v));
    else
        atomicMin((unsigned int *)addr, __float_as_uint(v));
}

// PTX griddepcontrol (sm_90+): no-ops when the kernel was launched without the programmatic-serialization attribute
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- TMA-staged raw IQ (STAGED kernels): the raw codes of a chunk arrive in shared memory by one bulk
// copy (UBLKCP, completion on an mbarrier) issued by a single thread two iterations ahead, instead of
// 16 two-byte LDGs per thread held in 16 registers for a whole iteration.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *mbar) {
    // the buffer was last read through the generic proxy (LDS) before the barrier this thread just passed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(mbar)), "r"(parity)
                     : "memory");
}
// the bulk copy thread 0 issues right after the first exchange barrier of an iteration (everybody has
// consumed the buffer it refills by then)
struct StageNext {
    void *dst = nullptr;
    const void *src = nullptr;
    uint32_t bytes = 0;
    unsigned long long *mbar = nullptr;
};

// Passes 1 .. PASSES-1 of one frame.  With two shared frames (NBUF == 2) pass k scatters into
// buffer (k-1)&1 and gathers from it after ONE barrier: the buffer being overwritten was last
// read two barriers ago.  With a single frame a second barrier protects the gather.
template <int NL, int S, int IN, int OUT, int PASS>
struct MiddlePasses {
    // tw_mid: tables of the middle passes (shared-memory copy when it fits), tw_all: the
    // complete global table (the last pass reads it when its twiddles are not in registers)
    static __device__ __forceinline__ void run(cf *x0, cf *x1, const cf *tw_mid, const cf *tw_all, const cf *twreg,
                                               int tid, cf *u, int tbase = 0, const StageNext *sn = nullptr) {
        using F = SpectrumFrame<NL, S, IN, OUT>;
        if constexpr (PASS < Plan<NL>::PASSES) {
            cf *x = ((PASS - 1) & 1) ? x1 : x0;
#ifdef RFA_TWPRE
            // twiddles of a middle pass are fetched BEFORE the exchange barrier (their addresses do not depend on the
            // exchange, and the 32 registers of u[] are dead between scatter and gather): the loads fly while the CTA waits
            constexpr bool PRE = PASS != F::LAST && F::MID_TW_SMEM && (Geom<NL>::E / Plan<NL>::radix(PASS)) * (Plan<NL>::radix(PASS) - 1) <= 16;
            cf twpre[PRE ? (Geom<NL>::E / Plan<NL>::radix(PASS)) * (Plan<NL>::radix(PASS) - 1) : 1];
#else
            constexpr bool PRE = false;
            cf twpre[1];
#endif
#ifndef RFA_EXP_NOXCHG
            F::template scatter<PASS - 1>(x, tid, u);
            if constexpr (PRE) F::template load_pass_tw<PASS>(tw_mid, tid, twpre);
            RFA_STAMP(tbase + 2 * PASS - 1);  // scatter issued
            __syncthreads();
            RFA_STAMP(tbase + 2 * PASS);      // barrier passed
            if (PASS == 1 && sn != nullptr && sn->bytes != 0 && threadIdx.x == 0) tma_load_1d(sn->dst, sn->src, sn->bytes, sn->mbar);
#endif
            if constexpr (PASS == F::LAST && F::LAST_TW_REG)
                F::gather_last_reg(x, twreg, tid, u);
            else if constexpr (PASS == F::LAST)
                F::template gather<PASS>(x, tw_all, tid, u);
            else if constexpr (PRE)
                F::template gather_reg<PASS>(x, twpre, tid, u);
            else
                F::template gather<PASS>(x, tw_mid, tid, u);
            if constexpr (Geom<NL>::NBUF == 1 && PASS + 1 < Plan<NL>::PASSES) __syncthreads();
            MiddlePasses<NL, S, IN, OUT, PASS + 1>::run(x0, x1, tw_mid, tw_all, twreg, tid, u, tbase);
        }
    }
};

// Frame schedule.  Work item v (0 .. nframes-1) is frame nframes-1-v: the newest frames are
// transformed FIRST.  Items are handed out in chunks of FPC consecutive items (one per frame slot
// of the CTA): every group starts on chunks `group` and `group + groups`, later chunks come from an
// atomic counter, fetched two iterations ahead so that the raw IQ of the next chunk is in flight
// while the current one is transformed.  (On B200 the two CTAs of an SM do not run at the same
// speed -- the scheduler favours one of them -- so a static split leaves the SM half empty at the
// end: gpurun_out/trace1.log.)
// The newest avg_len+1 frames ("tail") are finished in the first iteration or two and counted in a
// ticket; one extra CTA at the end of the grid computes the time average from them (average_cta).
// ticket[0..3]: finished tail rows per residue, [4]: CTAs done, [8..11]: chunk counters per
// residue; the last CTA of the launch re-arms them all for the next launch.
// Timing experiments (tools/): RFA_MINCTAS changes the occupancy target, the
// RFA_EXP_* switches drop one part of the kernel (results are then wrong).
#ifndef RFA_MINCTAS
#define RFA_MINCTAS 2
#endif
enum : int { TICKET_TAIL = 0, TICKET_DONE = 4, TICKET_WORK = 8, TICKET_WORDS = 12 };

// AnalyzerSurface.kt:710-714: avg[i] = (sum of the newest L+1 rows, newest -> oldest, float32) / (L+1).
// Run by the launch's LAST CTA, which does nothing else.  The launcher leaves it one resident slot
// (workers = capacity - 1), so it sleeps beside the workers until the tail rows are counted, a few
// microseconds into the launch, averages while they transform, and leaves; the dynamic chunk
// hand-out makes up for the missing worker.  No worker ever waits, and the transform loop carries
// none of this code.
__device__ __forceinline__ float average_row_term(const SpectrumParams &p, int r, long long i) {
    long long row = p.avg_newest + (long long)r * p.avg_dir;
    if (p.ring_rows > 0) {
        row %= p.ring_rows;
        if (row < 0) row += p.ring_rows;
    }
    return __ldcg(p.rows + row * p.row_stride + i);
}
__device__ __forceinline__ void average_cta(const SpectrumParams &p, int S, int N, int n_tail) {
    // where the newest avg_len+1 rows start: once per CTA (a 64-bit modulo per row and bin made this CTA the launch's
    // last one to finish at 8192 points on 256 threads -- 32 trips of nine dependent address chains and an L2 round trip)
    __shared__ long long s_row_off[31];
    if (threadIdx.x < 31) {
        long long row = p.avg_newest + (long long)threadIdx.x * p.avg_dir;
        if (p.ring_rows > 0) {
            row %= p.ring_rows;
            if (row < 0) row += p.ring_rows;
        }
        s_row_off[threadIdx.x] = row * p.row_stride;
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < S; c++) {
            volatile unsigned int *t = p.ticket + TICKET_TAIL + c;
            while (*t < (unsigned int)n_tail) __nanosleep(256);
        }
    }
    __syncthreads();
    __threadfence();  // the workers' rows, published before their counts
    const int nr = p.avg_len + 1 < 31 ? p.avg_len + 1 : 31;
    const float inv_terms = (float)(p.avg_len + 1);
    // two bins per trip: twice the loads in flight per L2 round trip
#pragma unroll 1
    for (int i0 = (int)threadIdx.x; i0 < N; i0 += 2 * (int)blockDim.x) {
        const int i1 = i0 + (int)blockDim.x;
        float v0[31], v1[31];
#pragma unroll
        for (int r = 0; r < 31; r++) {
            v0[r] = v1[r] = -9999.0f;
            if (r < nr && r < p.avg_valid) {
                const float *row = p.rows + s_row_off[r];
                v0[r] = __ldcg(row + i0);
                if (i1 < N) v1[r] = __ldcg(row + i1);
            }
        }
        float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
        for (int r = 0; r < 31; r++)
            if (r < nr) {
                sum0 = __fadd_rn(sum0, v0[r]);
                sum1 = __fadd_rn(sum1, v1[r]);
            }
        p.avg[i0] = __fdiv_rn(sum0, inv_terms);
        if (i1 < N) p.avg[i1] = __fdiv_rn(sum1, inv_terms);
    }
}

// every thread of a worker CTA that has just stored `cnt` tail rows of residue c (CTA-uniform)
__device__ __forceinline__ void publish_tail(const SpectrumParams &p, int c, int cnt) {
    __threadfence();  // rows of this CTA visible before the count
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p.ticket + TICKET_TAIL + c, (unsigned int)cnt);
}

// the last CTA of the launch to leave re-arms the counters for the next launch
__device__ __forceinline__ void retire_cta(const SpectrumParams &p) {
    if (threadIdx.x == 0) {
        if (atomicAdd(p.ticket + TICKET_DONE, 1u) == gridDim.x - 1u) {
#pragma unroll
            for (int k = 0; k < TICKET_WORDS; k++) p.ticket[k] = 0u;
        }
    }
}

// chunk buffers of a STAGED kernel: a two-deep ring, or ONE buffer where two no longer fit beside the exchange frames
// (int16 IQ at N >= 8192: the refill issued after the first exchange barrier then has the rest of the frame to land)
template <int NL, int IN>
RFA_CX int stage_buffers() {
    return (IN == FMT_S16LE && NL >= 8192) ? 1 : 2;
}
// bytes of shared memory a STAGED kernel needs on top of SpectrumFrame::SMEM_BYTES
template <int NL, int IN>
RFA_CX size_t staged_bytes() {
    return (size_t)stage_buffers<NL, IN>() * Geom<NL>::FPC * NL * in_elem_bytes<IN>();
}
template <int NL, int S, int IN, int OUT>
RFA_CX size_t staged_offset() {  // 128-byte aligned start of the chunk buffers
    return (SpectrumFrame<NL, S, IN, OUT>::SMEM_BYTES + 127) / 128 * 128;
}

// chunk q of a STAGED kernel: the block of valid frames it covers, as (first frame, bytes)
template <int NL, int FPC, int BPS>
__device__ __forceinline__ void chunk_block(long long nframes, int q, long long *f_start, uint32_t *bytes) {
    const long long f_hi = nframes - 1 - (long long)q * FPC;  // newest frame of the chunk (sub = 0)
    long long f_lo = f_hi - (FPC - 1);
    if (f_lo < 0) f_lo = 0;
    *f_start = f_lo;
    *bytes = f_hi >= f_lo ? (uint32_t)((f_hi - f_lo + 1) * (long long)NL * BPS) : 0u;
}

template <int NL, int S, int IN, int OUT, bool STAGED = false>
#ifdef RFA_R8
__global__ void __launch_bounds__(Geom<NL>::CTA, (Geom<NL>::CTA <= 256 || NL == 4096) ? RFA_MINCTAS : 1) spectrum_kernel(const SpectrumParams p) {
#else
__global__ void __launch_bounds__(Geom<NL>::CTA, Geom<NL>::CTA <= 256 ? RFA_MINCTAS : 1) spectrum_kernel(const SpectrumParams p) {
#endif
    using G = Geom<NL>;
    using F = SpectrumFrame<NL, S, IN, OUT>;
    constexpr int T = G::T, E = G::E, FPC = G::FPC, N = NL * S;
    static_assert(!STAGED || (S == 1 && (IN == FMT_S8 || IN == FMT_U8 || IN == FMT_S16LE)), "staging is for the integer formats");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_chunk[2];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    RFA_STAMP(0);
    const bool want_avg = (OUT == OUT_DB) && p.avg != nullptr;
    const bool want_peak = (OUT == OUT_DB) && p.peaks != nullptr;
    const int n_tail = want_avg ? (int)(p.nframes < p.avg_len + 1 ? p.nframes : p.avg_len + 1) : 0;
    const int nworkers = (int)gridDim.x - (want_avg ? 1 : 0);  // the launcher adds the averaging CTA
    // Programmatic dependent launch (p.pdl, kernels launched back to back): the next launch's CTAs take the
    // slots this grid's CTAs free, load their constant tables while this grid drains, and only then wait for it --
    // everything that touches IQ bytes, rows, peaks or the ticket counters sits behind grid_dependency_wait().
    const bool pdl = p.pdl != 0;
    if (pdl) launch_dependents();
    if ((int)blockIdx.x == nworkers) {
        if (pdl) grid_dependency_wait();
        average_cta(p, S, N, n_tail);
        retire_cta(p);
        return;
    }
    const int sub = threadIdx.x / T, tid = threadIdx.x % T;
    cf *x0 = reinterpret_cast<cf *>(smem_raw) + (size_t)sub * G::NBUF * Plan<NL>::SMEM_POINTS;
    cf *x1 = G::NBUF == 2 ? x0 + Plan<NL>::SMEM_POINTS : x0;

    const int c = (S == 1) ? 0 : (int)(blockIdx.x % S);
    const int groups = nworkers / S, group = blockIdx.x / S;
    const int nchunks = (int)((p.nframes + FPC - 1) / FPC);  // < 2^31, checked by the launcher
    constexpr int BPS = in_elem_bytes<IN>();

    // chunk q -> item v = q*FPC + sub -> frame nframes-1-v
    int q = group, q_next = group + groups, q_next2 = group + 2 * groups;
    cf u[E];
    uint32_t raw[F::PREFETCH ? E : 1];
    constexpr size_t CHUNK_BYTES = (size_t)FPC * NL * BPS;
    unsigned char *stage = smem_raw + staged_offset<NL, S, IN, OUT>();  // [2][CHUNK_BYTES] (STAGED)
    constexpr int NSTG = STAGED ? stage_buffers<NL, IN>() : 2;
    auto start_chunks = [&]() {  // thread 0: the first chunks start moving
        mbar_init(&s_mbar[0]);
        mbar_init(&s_mbar[1]);
        const int qs[2] = {q, q_next};
#pragma unroll
        for (int b = 0; b < NSTG; b++) {
            long long f0;
            uint32_t bytes;
            chunk_block<NL, FPC, BPS>(p.nframes, qs[b], &f0, &bytes);
            if (qs[b] < nchunks && bytes)
                tma_load_1d(stage + b * CHUNK_BYTES, (const char *)p.in + f0 * (long long)NL * BPS, bytes, &s_mbar[b]);
        }
    };
    if constexpr (STAGED) {  // before anything else -- unless the previous launch may still be running
        if (!pdl && threadIdx.x == 0) start_chunks();
    } else if constexpr (F::PREFETCH) {  // first: get the raw IQ of the first frame moving
        const long long v = (long long)q * FPC + sub;
        if (!pdl && v < p.nframes) F::load_raw((const char *)p.in + ((p.nframes - 1 - v) * (long long)N + tid) * BPS, raw);
    }

    // middle-pass twiddle tables: one shared-memory copy per CTA
    const cf *tw = p.tw;
    if constexpr (F::MID_TW_SMEM) {
        cf *stw = reinterpret_cast<cf *>(smem_raw + G::SMEM);
        for (int i = threadIdx.x; i < F::MID_TW; i += G::CTA) stw[i] = p.tw[i];
        tw = stw;
    }
    // last-pass twiddles and window taps of this thread never change: registers.
    // The window is pre-multiplied by the format's power-of-two unit.
    cf twreg[F::LAST_TW_REG ? F::LAST_TW : 1];
    if constexpr (F::LAST_TW_REG) F::load_last_tw(p.tw, tid, twreg);
    float wreg[E];
    if (S == 1) {
        constexpr int R = Plan<NL>::radix(0), STR = NL / R;
#pragma unroll
        for (int b = 0; b < E / R; b++)
#pragma unroll
            for (int r = 0; r < R; r++)
                wreg[b * R + r] = (p.win ? p.win[tid + b * T + r * STR] : 1.0f) * unit_scale<IN>();
    }
    float pk[E];
#pragma unroll
    for (int e = 0; e < E; e++) pk[e] = -999999.0f;
    const float inv_n2 = p.inv_n2;
    bool worked = false;
    unsigned int pending = 0;
    if (pdl) {
        grid_dependency_wait();
        if constexpr (STAGED) {
            if (threadIdx.x == 0) start_chunks();
        } else if constexpr (F::PREFETCH) {
            const long long v = (long long)q * FPC + sub;
            if (v < p.nframes) F::load_raw((const char *)p.in + ((p.nframes - 1 - v) * (long long)N + tid) * BPS, raw);
        }
    }
    if (threadIdx.x == 0) pending = atomicAdd(p.ticket + TICKET_WORK + c, 1u);
    __syncthreads();  // twiddle copy

    RFA_STAMP(1);  // prologue done (slot 0 is stamped at kernel entry)
    for (int it = 0; q < nchunks; it++) {  // q is CTA-uniform
        const long long v = (long long)q * FPC + sub;
        const bool active = v < p.nframes;
        const long long f = p.nframes - 1 - v;
        RFA_STAMP(8 + 8 * it);  // iteration start; +1..+4 exchanges, +6 after the last butterflies, +7 after emit
        // chunk after next: thread 0 asks now, everybody reads it at the end of the iteration
        // the chunk for three iterations on: thread 0 posts the counter value it asked for one iteration
        // ago (long since returned) and asks for the next; everybody reads the post at the end of the iteration
        if (threadIdx.x == 0) {
            s_chunk[it & 1] = 3 * groups + (int)pending;
            pending = atomicAdd(p.ticket + TICKET_WORK + c, 1u);
        }
        StageNext sn;
        if constexpr (STAGED) {
            // this chunk's raw IQ sits in buffer it & 1 (copy issued two iterations ago); the buffer is
            // refilled with chunk q_next2 right after the first exchange barrier of this iteration
            // (one buffer, NSTG == 1: copy issued one iteration ago, refilled with chunk q_next)
            long long f0;
            uint32_t bytes;
            chunk_block<NL, FPC, BPS>(p.nframes, q, &f0, &bytes);
            const int sb = NSTG == 2 ? (it & 1) : 0;
            mbar_wait(&s_mbar[sb], (uint32_t)((NSTG == 2 ? (it >> 1) : it) & 1));
            if (active) {
                if constexpr (F::PREFETCH) {
                    F::load_raw((const char *)(stage + sb * CHUNK_BYTES) + ((f - f0) * (long long)NL + tid) * BPS, raw);
                    F::first_from_raw(raw, wreg, u);
                } else {  // 32 points per thread: convert straight out of the staged chunk
                    F::template first<true>(p, f, c, tid, wreg, u, (const char *)(stage + sb * CHUNK_BYTES) + (f - f0) * (long long)NL * BPS);
                }
            }
            const int q_refill = NSTG == 2 ? q_next2 : q_next;
            if (q_refill < nchunks) {
                chunk_block<NL, FPC, BPS>(p.nframes, q_refill, &f0, &sn.bytes);
                sn.dst = stage + sb * CHUNK_BYTES;
                sn.src = (const char *)p.in + f0 * (long long)NL * BPS;
                sn.mbar = &s_mbar[sb];
            }
        } else if constexpr (F::PREFETCH) {
            if (active) F::first_from_raw(raw, wreg, u);
            const long long vn = (long long)q_next * FPC + sub;  // lands while this frame is transformed
            if (vn < p.nframes) F::load_raw((const char *)p.in + ((p.nframes - 1 - vn) * (long long)N + tid) * BPS, raw);
        } else {
            if (active) F::first(p, f, c, tid, wreg, u);
        }
        // single frame buffer: the previous frame's last gather must be done before the first
        // scatter; with ping-pong buffers and an even number of exchanges the barriers inside
        // the passes already order it (PASSES-1 exchanges alternate A,B,A,...)
        if (it > 0 && (G::NBUF == 1 || ((Plan<NL>::PASSES - 1) & 1))) __syncthreads();
        if constexpr (Plan<NL>::PASSES > 1)
            MiddlePasses<NL, S, IN, OUT, 1>::run(x0, x1, tw, p.tw, twreg, tid, u, 8 + 8 * it, STAGED ? &sn : nullptr);
        else
            __syncthreads();
        RFA_STAMP(8 + 8 * it + 6);
        if (active) {
            worked = true;
            float *out = p.rows + frame_row(p, f) * p.row_stride;
            if (f >= p.store_from) {
                if (want_peak)
                    F::template emit<true, true>(out, c, tid, u, pk, inv_n2);
                else
                    F::template emit<false, true>(out, c, tid, u, pk, inv_n2);
            } else if (want_peak) {
                F::template emit<true, false>(out, c, tid, u, pk, inv_n2);
            }
        }
        RFA_STAMP(8 + 8 * it + 7);
        // tail rows of this chunk (CTA-uniform test; only the first iteration or two)
        if (q < 32 && q * FPC < n_tail) {  // n_tail <= 31
            const int cnt = (n_tail - q * FPC < FPC) ? n_tail - q * FPC : FPC;
            publish_tail(p, c, cnt);
        }
        q = q_next;
        q_next = q_next2;
        q_next2 = s_chunk[it & 1];  // written before a barrier of this iteration, rewritten two iterations on
    }

    RFA_STAMP(2);  // main loop done
    if constexpr (FPC > 1) {
        // several frame slots per CTA hold the same bins: reduce them in shared memory first, then
        // one atomic per bin per CTA (at N = 256 the 16 slots otherwise queue 4736 atomics per bin)
        if (want_peak) {
            float *red = reinterpret_cast<float *>(smem_raw);  // [FPC][NL], the exchange frames are free now
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; e++) red[sub * NL + F::peak_index(c, tid, e)] = worked ? pk[e] : -999999.0f;
            __syncthreads();
            for (int i = threadIdx.x; i < NL; i += G::CTA) {
                float m = red[i];
#pragma unroll
                for (int k = 1; k < FPC; k++) m = fmaxf(m, red[k * NL + i]);
                if (m > -999999.0f) atomic_max_float(p.peaks + i, m);
            }
        }
    } else if (want_peak && worked) {
#pragma unroll
        for (int e = 0; e < E; e++) atomic_max_float(p.peaks + F::peak_index(c, tid, e), pk[e]);
    }
    RFA_STAMP(3);  // peak atomics issued
    retire_cta(p);
    RFA_STAMP(4);  // kernel end
#ifdef RFA_TRACE
    if (threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_trace[blockIdx.x & 511][5] = smid;
    }
#endif
}
#endif  // __CUDACC__

}  // namespace rfa
