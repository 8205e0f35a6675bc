// spectrum_lean_kernel.cuh -- three-CTAs-per-SM variant of the fused IQ -> spectrum kernel for N = 4096, 8-bit IQ
// (RFA_LEAN=1).
//
// Same computation, reference lines and per-thread phase functions as spectrum_kernel.cuh.  The default kernel keeps
// window taps (16), last-pass twiddles (30) and running peaks (16) in registers: 128 registers, two 256-thread CTAs
// per SM, four warps per scheduler -- and reaches 62 % of its issue-slot bound.  Here the window taps
// (pre-multiplied by the format's unit, one copy per CTA in shared memory) and the last-pass twiddles (global table,
// L1-resident: every CTA of the SM reads the same 30 KB) are fetched when they are used -- LDS/LDG issue in the shadow
// of the packed FP32 instructions (profiles/r01b_ubench_coissue.txt) -- so that a thread fits 80 registers and THREE
// CTAs share an SM: six warps per scheduler in three independent phase groups.  One exchange frame per CTA (69 KB of
// shared memory per CTA), i.e. two barriers per exchange instead of one.
#pragma once
#include "spectrum_kernel.cuh"

namespace rfa {

struct GeomLean {
    static constexpr int NL = 4096, T = 256, CTA = 256;
    static constexpr size_t XCHG = (size_t)Plan<NL>::SMEM_POINTS * sizeof(cf);
    static constexpr int MID_TW = pass_tw_offset<NL>(2);
    static constexpr size_t TW_OFF = XCHG;
    static constexpr size_t WIN_OFF = (TW_OFF + MID_TW * sizeof(cf) + 127) / 128 * 128;
    static constexpr size_t STAGE_OFF = WIN_OFF + (size_t)NL * sizeof(float);
    static RFA_CX size_t smem(int bps) { return STAGE_OFF + 2 * (size_t)NL * bps; }
};

#ifdef __CUDACC__
template <int IN>
__global__ void __launch_bounds__(GeomLean::CTA, 3) spectrum_lean_kernel(const SpectrumParams p) {
    using G = GeomLean;
    using F = SpectrumFrame<G::NL, 1, IN, OUT_DB>;
    constexpr int NL = G::NL, T = G::T, E = 16;
    constexpr int BPS = in_elem_bytes<IN>();
    constexpr size_t CHUNK_BYTES = (size_t)NL * BPS;
    static_assert(IN == FMT_S8 || IN == FMT_U8, "lean variant: 8-bit IQ (two 16-bit chunk buffers would cost the third CTA)");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_chunk[2];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    const bool want_avg = p.avg != nullptr, want_peak = p.peaks != nullptr;
    const int n_tail = want_avg ? (int)(p.nframes < p.avg_len + 1 ? p.nframes : p.avg_len + 1) : 0;
    const int nworkers = (int)gridDim.x - (want_avg ? 1 : 0);  // the launcher adds the averaging CTA
    if ((int)blockIdx.x == nworkers) {
        average_cta(p, 1, NL, n_tail);
        retire_cta(p);
        return;
    }
    const int tid = threadIdx.x;
    cf *x = reinterpret_cast<cf *>(smem_raw);
    cf *stw = reinterpret_cast<cf *>(smem_raw + G::TW_OFF);
    float *swin = reinterpret_cast<float *>(smem_raw + G::WIN_OFF);
    unsigned char *stage = smem_raw + G::STAGE_OFF;  // [2][CHUNK_BYTES]
    const int nchunks = (int)p.nframes;              // one frame per chunk; < 2^31, checked by the launcher
    const int groups = nworkers, group = (int)blockIdx.x;
    // frames are handed out newest-first: three static chunks per CTA, then an atomic counter read three iterations ahead
    int q = group, q_next = group + groups, q_next2 = group + 2 * groups;
    auto issue = [&](int qq, int buf) {  // thread 0 only
        if (qq < nchunks)
            tma_load_1d(stage + (size_t)buf * CHUNK_BYTES, (const char *)p.in + (p.nframes - 1 - qq) * (long long)NL * BPS,
                        (uint32_t)CHUNK_BYTES, &s_mbar[buf]);
    };
    unsigned int pending = 0;
    if (tid == 0) {
        mbar_init(&s_mbar[0]);
        mbar_init(&s_mbar[1]);
        issue(q, 0);
        issue(q_next, 1);
        pending = atomicAdd(p.ticket + TICKET_WORK, 1u);
    }
    for (int i = tid; i < G::MID_TW; i += G::CTA) stw[i] = p.tw[i];
    for (int i = tid; i < NL; i += G::CTA) swin[i] = (p.win ? p.win[i] : 1.0f) * unit_scale<IN>();
    float pk[E];
#pragma unroll
    for (int e = 0; e < E; e++) pk[e] = -999999.0f;
    const float inv_n2 = p.inv_n2;
    bool worked = false;
    __syncthreads();  // tables, mbarrier init

    for (int it = 0; q < nchunks; it++) {  // q is CTA-uniform
        const long long f = p.nframes - 1 - q;
        if (tid == 0) {
            s_chunk[it & 1] = 3 * groups + (int)pending;
            pending = atomicAdd(p.ticket + TICKET_WORK, 1u);
        }
        cf u[E];
        mbar_wait(&s_mbar[it & 1], (uint32_t)((it >> 1) & 1));
        {
            const char *src = (const char *)(stage + (size_t)(it & 1) * CHUNK_BYTES) + (size_t)tid * BPS;
#pragma unroll
            for (int r = 0; r < E; r++)
                u[r] = decode_point<IN>((uint32_t)((const uint16_t *)src)[r * T], swin[tid + r * T]);
            Dft<16>::run(u);
        }
        if (it > 0) __syncthreads();  // the previous frame's last gather is done with the exchange frame
        F::template scatter<0>(x, tid, u);
        __syncthreads();
        if (tid == 0) issue(q_next2, it & 1);  // everybody has consumed this chunk's raw codes
        F::template gather<1>(x, stw, tid, u);
        __syncthreads();
        F::template scatter<1>(x, tid, u);
        __syncthreads();
        F::template gather<2>(x, p.tw, tid, u);  // last-pass twiddles through L1
        {
            float *out = p.rows + frame_row(p, f) * p.row_stride;
            if (f >= p.store_from) {
                if (want_peak)
                    F::template emit<true, true>(out, 0, tid, u, pk, inv_n2);
                else
                    F::template emit<false, true>(out, 0, tid, u, pk, inv_n2);
            } else if (want_peak) {
                F::template emit<true, false>(out, 0, tid, u, pk, inv_n2);
            }
            worked = true;
        }
        if (q < n_tail) publish_tail(p, 0, 1);  // CTA-uniform; only the newest avg_len+1 frames
        q = q_next;
        q_next = q_next2;
        q_next2 = s_chunk[it & 1];  // written before a barrier of this iteration, rewritten two iterations on
    }
    if (want_peak && worked) {
#pragma unroll
        for (int e = 0; e < E; e++) atomic_max_float(p.peaks + F::peak_index(0, tid, e), pk[e]);
    }
    retire_cta(p);
}
#endif  // __CUDACC__

}  // namespace rfa
