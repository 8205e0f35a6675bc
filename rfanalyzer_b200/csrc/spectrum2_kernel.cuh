// spectrum2_kernel.cuh -- dual-frame variant of the fused IQ -> spectrum kernel for N = 256 .. 4096.
//
// Same computation and the same reference lines as spectrum_kernel.cuh (convert, window, FFT, dB,
// fft-shift, row store, peak hold, time average); what changes is how the work sits on the SM.
// Measurements of the single-frame kernel on B200 showed that at 128
// registers per thread ptxas has no room to hoist shared-memory loads, that the per-thread
// invariants (window taps, last-pass twiddles, running peaks: 62 registers) are paid once per
// resident frame, and that 8 warps with a free register budget run as fast as 16 warps at 128.
// So here one thread carries the SAME 16 points of TWO consecutive frames (A = older, B = newer)
// as c2 values: every FP32 operation is a packed FADD2/FMUL2/FFMA2 over both frames, window taps,
// twiddles and the peak registers are shared by the pair, shared-memory exchanges move 16 bytes
// per instruction, and one barrier serves two frames.  One 256-thread CTA per SM, up to 255
// registers per thread.
#pragma once
#include "spectrum_kernel.cuh"

namespace rfa {

template <int NL>
struct Geom2 {
    static constexpr int T = NL / 16;    // threads per frame pair
    static constexpr int E = 16;         // points of each frame per thread
    static constexpr int CTA = 256;
    static constexpr int FPC = CTA / T;  // frame pairs per CTA
    static constexpr size_t SMEM = (size_t)FPC * 2 * Plan<NL>::SMEM_POINTS * sizeof(c2);  // ping-pong frames
    static_assert(NL >= 256 && NL <= 4096, "dual-frame geometry covers 256 .. 4096 points");
};

// raw codes of the same point of frames A and B -> c2 * ws (see decode_point)
template <int IN>
RFA_HD c2 decode_pair(uint32_t a, uint32_t b, float ws) {
    if (IN == FMT_S8) {
        a ^= 0x8080u;
        b ^= 0x8080u;
        const c2 m = c2{magic_byte0(a), magic_byte0(b), magic_byte1(a), magic_byte1(b)};
        return cscale(cadd(m, c2{-8388736.0f, -8388736.0f, -8388736.0f, -8388736.0f}), ws);
    } else if (IN == FMT_U8) {
        const c2 m = c2{magic_byte0(a), magic_byte0(b), magic_byte1(a), magic_byte1(b)};
        const c2 k = cadd(m, c2{-8388608.0f, -8388608.0f, -8388608.0f, -8388608.0f});
        return cscale(cadd(k, c2{-127.4f, -127.4f, -127.4f, -127.4f}), ws);
    } else {
        a ^= 0x80008000u;
        b ^= 0x80008000u;
        const c2 m = c2{magic_half0(a), magic_half0(b), magic_half1(a), magic_half1(b)};
        return cscale(cadd(m, c2{-8421376.0f, -8421376.0f, -8421376.0f, -8421376.0f}), ws);
    }
}

// dB of both frames' bin: 1.50515*log2(re^2 + im^2) + bias (see logmag_db)
RFA_HD cf logmag_db2(c2 v, float db_bias) {
#ifdef RFA_PACKED
    const cf pw = fma2(cpk(v.rA, v.rB), cpk(v.rA, v.rB), RFA_PK(mul2(cpk(v.iA, v.iB), cpk(v.iA, v.iB))));
    float la, lb;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(pw.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(pw.y));
    return fma2(cpk(la, lb), cpk(1.5051499783199060f, 1.5051499783199060f), cpk(db_bias, db_bias));
#else
    return cf{logmag_db(cf{v.rA, v.iA}, db_bias), logmag_db(cf{v.rB, v.iB}, db_bias)};
#endif
}

template <int NL, int IN>
struct SpectrumFrame2 {
    using G = Geom2<NL>;
    using PL = Plan<NL>;
    static constexpr int T = G::T, E = G::E, N = NL;
    static constexpr int LAST = PL::PASSES - 1;
    static_assert(IN == FMT_S8 || IN == FMT_U8 || IN == FMT_S16LE, "integer IQ formats only");
    static_assert(PL::PASSES >= 2, "at least one exchange");
    static constexpr int LAST_TW = (E / PL::radix(LAST)) * (PL::radix(LAST) - 1);  // <= 15 complex
    static constexpr int MID_TW = PL::PASSES > 2 ? pass_tw_offset<NL>(LAST) : 0;
    static constexpr size_t SMEM_BYTES = G::SMEM + MID_TW * sizeof(cf);

    // `src` points at point `tid` of the frame
    static RFA_HD void load_raw(const char *src, uint32_t *raw) {
        constexpr int R = PL::radix(0), STR = NL / R;
        static_assert(R == E, "first pass is one radix-16 butterfly per thread");
#pragma unroll
        for (int r = 0; r < R; r++)
            raw[r] = (IN == FMT_S16LE) ? ((const uint32_t *)src)[r * STR] : (uint32_t)((const uint16_t *)src)[r * STR];
    }
    static RFA_HD void first_from_raw(const uint32_t *rawA, const uint32_t *rawB, const float *wreg, c2 *u) {
#pragma unroll
        for (int e = 0; e < E; e++) u[e] = decode_pair<IN>(rawA[e], rawB[e], wreg[e]);
        pass_first_compute<NL, T, PL::radix(0)>(u);
    }
    template <int PASS>
    static RFA_HD void scatter(c2 *x, int tid, const c2 *u) {
        pass_scatter<NL, T, PL::radix(PASS), PL::prod(PASS)>(x, tid, u);
    }
    template <int PASS>
    static RFA_HD void gather(const c2 *x, const cf *tw, int tid, c2 *u) {
        pass_gather<NL, T, PL::radix(PASS), PL::prod(PASS)>(x, tw + pass_tw_offset<NL>(PASS), tid, u);
    }
    static RFA_HD void load_last_tw(const cf *tw, int tid, cf *twreg) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST), NB = E / R;
        const cf *t = tw + pass_tw_offset<NL>(LAST);
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int r = 1; r < R; r++) twreg[b * (R - 1) + r - 1] = t[(r - 1) * P + ((tid + b * T) & (P - 1))];
    }
    static RFA_HD void gather_last_reg(const c2 *x, const cf *twreg, int tid, c2 *u) {
        constexpr int R = PL::radix(LAST), NB = E / R, STR = NL / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const c2 *xi = x + phys(tid + b * T);
#pragma unroll
            for (int r = 0; r < R; r++) {
                c2 v = xi[r * (STR + STR / 16)];
                if (r > 0) v = cmul(v, twreg[b * (R - 1) + r - 1]);
                u[b * R + r] = v;
            }
            Dft<R>::run(u + b * R);
        }
    }
    static RFA_CX int shifted_offset(int d) { return (PL::prod(LAST) <= N / 2) ? (d ^ (N >> 1)) : d; }
    // dB rows of both frames and the shared running peak; outA/outB are the row bases
    template <bool PEAK>
    static RFA_HD void emit(float *outA, float *outB, bool storeA, bool storeB, int tid, const c2 *u, float *pk,
                            float db_bias) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST), NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            float *oA = outA + (tid + b * T), *oB = outB + (tid + b * T);
#pragma unroll
            for (int cc = 0; cc < R; cc++) {
                const cf db = logmag_db2(u[b * R + Dft<R>::perm(cc)], db_bias);
                if (storeA) oA[shifted_offset(cc * P)] = db.x;
                if (storeB) oB[shifted_offset(cc * P)] = db.y;
                if (PEAK) pk[b * R + cc] = fmaxf(pk[b * R + cc], fmaxf(db.x, db.y));
            }
        }
    }
    static RFA_HD int peak_index(int tid, int e) {
        constexpr int R = PL::radix(LAST), P = PL::prod(LAST);
        return (tid + (e / R) * T + (e % R) * P) ^ (N >> 1);
    }
};

#ifdef __CUDACC__
template <int NL, int IN, int PASS>
struct MiddlePasses2 {
    static __device__ __forceinline__ void run(c2 *x0, c2 *x1, const cf *tw_mid, const cf *twreg, int tid, c2 *u,
                                               int tbase = 0) {
        using F = SpectrumFrame2<NL, IN>;
        if constexpr (PASS < Plan<NL>::PASSES) {
            c2 *x = ((PASS - 1) & 1) ? x1 : x0;
            F::template scatter<PASS - 1>(x, tid, u);
            RFA_STAMP(tbase + 2 * PASS - 1);
            __syncthreads();
            RFA_STAMP(tbase + 2 * PASS);
            if constexpr (PASS == F::LAST)
                F::gather_last_reg(x, twreg, tid, u);
            else
                F::template gather<PASS>(x, tw_mid, tid, u);
            MiddlePasses2<NL, IN, PASS + 1>::run(x0, x1, tw_mid, twreg, tid, u, tbase);
        }
    }
};

// Work item v (0 .. npairs-1) is the frame pair (A, B) = (nframes-2-2v, nframes-1-2v): newest pairs
// first, handed out in chunks of FPC pairs exactly like spectrum_kernel (two static chunks, then an
// atomic counter fetched ahead), tail rows counted per chunk and averaged by the extra last CTA.  With an odd frame count the last item has no frame A: it then
// transforms B twice and stores it once.
template <int NL, int IN>
__global__ void __launch_bounds__(256, 1) spectrum2_kernel(const SpectrumParams p) {
    using G = Geom2<NL>;
    using F = SpectrumFrame2<NL, IN>;
    constexpr int T = G::T, E = G::E, FPC = G::FPC, N = NL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_chunk[2];
    RFA_STAMP(0);
    const bool want_avg = p.avg != nullptr, want_peak = p.peaks != nullptr;
    const int n_tail = want_avg ? (int)(p.nframes < p.avg_len + 1 ? p.nframes : p.avg_len + 1) : 0;
    const int groups = (int)gridDim.x - (want_avg ? 1 : 0), group = blockIdx.x;  // + the averaging CTA
    if (group == groups) {
        average_cta(p, 1, N, n_tail);
        retire_cta(p);
        return;
    }
    const int sub = threadIdx.x / T, tid = threadIdx.x % T;
    c2 *x0 = reinterpret_cast<c2 *>(smem_raw) + (size_t)sub * 2 * Plan<NL>::SMEM_POINTS;
    c2 *x1 = x0 + Plan<NL>::SMEM_POINTS;

    const long long npairs = (p.nframes + 1) / 2;
    const int nchunks = (int)((npairs + FPC - 1) / FPC);
    constexpr int BPS = in_elem_bytes<IN>();
    const long long frame_bytes = (long long)N * BPS;

    // chunk q -> pair v = q*FPC + sub -> frames B = nframes-1-2v, A = B-1 (B twice when A < 0)
    int q = group, q_next = group + groups, q_next2 = group + 2 * groups;
    c2 u[E];
    uint32_t rawA[E], rawB[E];
    {
        const long long fB = p.nframes - 1 - 2 * ((long long)q * FPC + sub);
        if (fB >= 0) {
            const char *srcB = (const char *)p.in + (fB * (long long)N + tid) * BPS;
            F::load_raw(srcB, rawB);
            F::load_raw(fB > 0 ? srcB - frame_bytes : srcB, rawA);
        }
    }
    const cf *tw = p.tw;
    if constexpr (F::MID_TW > 0) {
        cf *stw = reinterpret_cast<cf *>(smem_raw + G::SMEM);
        for (int i = threadIdx.x; i < F::MID_TW; i += G::CTA) stw[i] = p.tw[i];
        tw = stw;
    }
    cf twreg[F::LAST_TW];
    F::load_last_tw(p.tw, tid, twreg);
    float wreg[E];
    {
        constexpr int STR = NL / 16;
#pragma unroll
        for (int r = 0; r < E; r++) wreg[r] = (p.win ? p.win[tid + r * STR] : 1.0f) * unit_scale<IN>();
    }
    float pk[E];
#pragma unroll
    for (int e = 0; e < E; e++) pk[e] = -999999.0f;
    const float db_bias = p.inv_n2;
    bool worked = false;
    unsigned int pending = 0;
    if (threadIdx.x == 0) pending = atomicAdd(p.ticket + TICKET_WORK + 0, 1u);
    __syncthreads();  // twiddle copy

    RFA_STAMP(1);
    for (int it = 0; q < nchunks; it++) {  // q is CTA-uniform
        const long long fB = p.nframes - 1 - 2 * ((long long)q * FPC + sub), fA = fB - 1;
        const bool active = fB >= 0;
        RFA_STAMP(8 + 8 * it);
        if (threadIdx.x == 0) {  // see spectrum_kernel: chunk for three iterations on
            s_chunk[it & 1] = 3 * groups + (int)pending;
            pending = atomicAdd(p.ticket + TICKET_WORK, 1u);
        }
        if (active) F::first_from_raw(rawA, rawB, wreg, u);
        {
            const long long fBn = p.nframes - 1 - 2 * ((long long)q_next * FPC + sub);  // lands while this pair is transformed
            if (fBn >= 0) {
                const char *srcB = (const char *)p.in + (fBn * (long long)N + tid) * BPS;
                F::load_raw(srcB, rawB);
                F::load_raw(fBn > 0 ? srcB - frame_bytes : srcB, rawA);
            }
        }
        if (it > 0 && ((Plan<NL>::PASSES - 1) & 1)) __syncthreads();
        MiddlePasses2<NL, IN, 1>::run(x0, x1, tw, twreg, tid, u, 8 + 8 * it);
        RFA_STAMP(8 + 8 * it + 6);
        if (active) {
            worked = true;
            const bool storeB = fB >= p.store_from, storeA = fA >= 0 && fA >= p.store_from;
            float *outB = p.rows + frame_row(p, fB) * p.row_stride;
            float *outA = p.rows + frame_row(p, fA >= 0 ? fA : fB) * p.row_stride;
            if (want_peak)
                F::template emit<true>(outA, outB, storeA, storeB, tid, u, pk, db_bias);
            else
                F::template emit<false>(outA, outB, storeA, storeB, tid, u, pk, db_bias);
        }
        RFA_STAMP(8 + 8 * it + 7);
        // this chunk's frames have ranks 2*q*FPC .. 2*(q+1)*FPC-1 (rank = nframes-1-frame); tail = ranks < n_tail
        if (q < 16 && 2 * q * FPC < n_tail) {  // n_tail <= 31
            const int cnt = (n_tail - 2 * q * FPC < 2 * FPC) ? n_tail - 2 * q * FPC : 2 * FPC;
            publish_tail(p, 0, cnt);
        }
        q = q_next;
        q_next = q_next2;
        q_next2 = s_chunk[it & 1];
    }

    RFA_STAMP(2);
    if constexpr (FPC > 1) {  // see spectrum_kernel: one atomic per bin per CTA
        if (want_peak) {
            float *red = reinterpret_cast<float *>(smem_raw);
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; e++) red[sub * NL + F::peak_index(tid, e)] = worked ? pk[e] : -999999.0f;
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
        for (int e = 0; e < E; e++) atomic_max_float(p.peaks + F::peak_index(tid, e), pk[e]);
    }
    RFA_STAMP(3);
    retire_cta(p);
    RFA_STAMP(4);
}
#endif  // __CUDACC__

}  // namespace rfa
