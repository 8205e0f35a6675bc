// spectrum_pair_kernel.cuh -- anti-phase variant of the fused IQ -> spectrum kernel for N = 4096 (RFA_PAIR=1).
//
// Same computation, reference lines and per-thread phase functions as spectrum_kernel.cuh (SpectrumFrame<4096,1,IN,
// OUT_DB>: conversion, window, three radix-16 passes, dB, fft-shift, row store, peak hold, time average by the
// launch's last CTA).  What changes is WHEN a frame's phases run.  The default kernel keeps two independent
// 256-thread CTAs per SM and measures FP32-pipe time + shared-memory-pipe time with almost no overlap (DESIGN.md
// 4.1): inside a CTA every warp is in the same phase, and the two CTAs drift through all relative phases.
// Here ONE 512-thread CTA per SM holds two frames ("slots" of 256 threads) that are kept exactly one segment
// apart by CTA-wide barriers, so that one slot's butterflies always run beside the other slot's exchange:
//
//     slot 0:  C20   X1    C1    X2   | C20   X1   ...        C20 = last pass + dB + store of the previous frame,
//     slot 1:  --    C20   X1    C1   | X2    C20  ...              raw codes + conversion + first pass of the next
//              ^ every '|' and every column boundary is a CTA barrier (4 per frame pair)       X1, X2 = scatter,
//                                                                    slot barrier (256 threads), gather;  C1 = middle pass
//
// The exchange inside a slot is ordered by a named barrier of the slot's 256 threads; one exchange frame per slot
// suffices (the CTA barrier after a gather orders it against the next scatter).  Raw IQ arrives by one bulk copy
// per chunk of two frames (two-deep ring) like in the STAGED default kernel; chunks are dealt round-robin.
#pragma once
#include "spectrum_kernel.cuh"

namespace rfa {

struct GeomPair {
    static constexpr int NL = 4096, T = 256, SLOTS = 2, CTA = T * SLOTS;
    static constexpr size_t XCHG = (size_t)SLOTS * Plan<NL>::SMEM_POINTS * sizeof(cf);  // one exchange frame per slot
    static constexpr int MID_TW = pass_tw_offset<NL>(2);                                // middle-pass table (240 entries)
    static constexpr size_t TW_OFF = XCHG;
    static constexpr size_t STAGE_OFF = (TW_OFF + MID_TW * sizeof(cf) + 127) / 128 * 128;
    static RFA_CX size_t chunk_bytes(int bps) { return (size_t)SLOTS * NL * bps; }
    static RFA_CX size_t smem(int bps) { return STAGE_OFF + 2 * chunk_bytes(bps); }
    static_assert(XCHG >= (size_t)SLOTS * NL * sizeof(float), "the peak reduction reuses the exchange frames");
};

#ifdef __CUDACC__
__device__ __forceinline__ void pair_slot_barrier(int sub) { asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory"); }
// CTA barrier as an arrival count on barrier 0: the two slots reach it from different places of their loops
// (slot 1 is one segment behind), which is what __syncthreads() formally does not allow
__device__ __forceinline__ void pair_cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// RFA_TRACE builds: thread 0 of each slot stamps clock64() around the CTA barriers of iterations 1..6
#ifdef RFA_TRACE
#define RFA_PSTAMP(it, k)                                                                                                   \
    do {                                                                                                                    \
        if (tid == 0 && (it) >= 1 && (it) <= 6) g_trace[blockIdx.x & 511][sub * 64 + ((it)-1) * 8 + (k)] = clock64();        \
    } while (0)
#else
#define RFA_PSTAMP(it, k) \
    do {                  \
    } while (0)
#endif

template <int IN>
__global__ void __launch_bounds__(GeomPair::CTA, 1) spectrum_pair_kernel(const SpectrumParams p) {
    using G = GeomPair;
    using F = SpectrumFrame<G::NL, 1, IN, OUT_DB>;
    constexpr int NL = G::NL, T = G::T, E = 16, FPC = G::SLOTS;
    constexpr int BPS = in_elem_bytes<IN>();
    constexpr size_t CHUNK_BYTES = G::chunk_bytes(BPS);
    constexpr int XSTR = NL / 16 + NL / 256;  // 272: gather stride of the radix-16 passes in the padded frame
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    const bool want_avg = p.avg != nullptr, want_peak = p.peaks != nullptr;
    const int n_tail = want_avg ? (int)(p.nframes < p.avg_len + 1 ? p.nframes : p.avg_len + 1) : 0;
    const int nworkers = (int)gridDim.x - (want_avg ? 1 : 0);  // the launcher adds the averaging CTA
    if ((int)blockIdx.x == nworkers) {
        average_cta(p, 1, NL, n_tail);
        retire_cta(p);
        return;
    }
    const int sub = threadIdx.x / T, tid = threadIdx.x % T;
    cf *x = reinterpret_cast<cf *>(smem_raw) + (size_t)sub * Plan<NL>::SMEM_POINTS;
    cf *stw = reinterpret_cast<cf *>(smem_raw + G::TW_OFF);
    unsigned char *stage = smem_raw + G::STAGE_OFF;  // [2][CHUNK_BYTES]
    const int nchunks = (int)((p.nframes + FPC - 1) / FPC);  // < 2^31, checked by the launcher
    const int stride = nworkers;
    int q = (int)blockIdx.x;

    auto issue = [&](int qq, int buf) {  // one thread
        long long f0;
        uint32_t bytes;
        chunk_block<NL, FPC, BPS>(p.nframes, qq, &f0, &bytes);
        if (qq < nchunks && bytes)
            tma_load_1d(stage + (size_t)buf * CHUNK_BYTES, (const char *)p.in + f0 * (long long)NL * BPS, bytes, &s_mbar[buf]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&s_mbar[0]);
        mbar_init(&s_mbar[1]);
        issue(q, 0);
        issue(q + stride, 1);
    }
    for (int i = threadIdx.x; i < G::MID_TW; i += G::CTA) stw[i] = p.tw[i];
    cf twreg[F::LAST_TW];
    F::load_last_tw(p.tw, tid, twreg);
    float wreg[E];
#pragma unroll
    for (int r = 0; r < E; r++) wreg[r] = (p.win ? p.win[tid + r * T] : 1.0f) * unit_scale<IN>();
    float pk[E];
#pragma unroll
    for (int e = 0; e < E; e++) pk[e] = -999999.0f;
    const float inv_n2 = p.inv_n2;
    const cf *twrow = stw + (tid >> 4);  // W_256^(c * (tid >> 4)) at twrow[(c - 1) * 16]
    const cf *xi = x + phys(tid);
    cf u[E];
    bool worked = false, prev_active = false;
    long long prev_f = 0;
    __syncthreads();  // twiddle copy, mbarrier init

    // last pass + dB + store + peak of the frame whose gathered points sit in u
    auto finish = [&](long long f) {
#pragma unroll
        for (int r = 1; r < E; r++) u[r] = cmul(u[r], twreg[r - 1]);
        Dft<16>::run(u);
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
    };
    // rows of the newest avg_len+1 frames are counted for the averaging CTA (slot-uniform test)
    auto publish = [&](long long f) {
        if (p.nframes - 1 - f < n_tail) {
            __threadfence();
            pair_slot_barrier(sub);
            if (tid == 0) atomicAdd(p.ticket + TICKET_TAIL, 1u);
        }
    };

    if (sub == 1) pair_cta_barrier();  // slot 1 runs one segment behind slot 0
    for (int it = 0; q < nchunks; it++, q += stride) {  // q, it are CTA-uniform
        const long long v = (long long)q * FPC + sub;
        const bool active = v < p.nframes;
        const long long f = p.nframes - 1 - v;
        // ---- C20: finish the previous frame, start this one
        RFA_PSTAMP(it, 0);
        if (prev_active) {
            finish(prev_f);
            publish(prev_f);
        }
        {
            long long f0;
            uint32_t bytes;
            chunk_block<NL, FPC, BPS>(p.nframes, q, &f0, &bytes);
            mbar_wait(&s_mbar[it & 1], (uint32_t)((it >> 1) & 1));
            if (active) {
                // the middle pass's Stockham twiddles are applied HERE, to the first pass's outputs (natural output c of
                // thread tid is input r = tid >> 4 of middle-pass butterfly k = c: W_256^(k r), the same table entry and
                // the same product as in spectrum_kernel's gather), so that the middle segment C1 loads nothing while the
                // other slot's exchange owns the shared-memory pipe; these loads issue after finish(), when it is idle
                uint32_t raw[E];
                cf twm[E - 1];
                F::load_raw((const char *)(stage + (size_t)(it & 1) * CHUNK_BYTES) + ((f - f0) * (long long)NL + tid) * BPS, raw);
#pragma unroll
                for (int c = 1; c < E; c++) twm[c - 1] = twrow[(c - 1) * 16];
                F::first_from_raw(raw, wreg, u);
#pragma unroll
                for (int c = 1; c < E; c++) u[Dft<16>::perm(c)] = cmul(u[Dft<16>::perm(c)], twm[c - 1]);
            }
        }
        RFA_PSTAMP(it, 1);
        pair_cta_barrier();
        RFA_PSTAMP(it, 2);
        // both slots have read this chunk's raw codes once slot 1 is past this barrier: its first thread refills the
        // buffer with the chunk two iterations on (an exchange segment has issue slots to spare)
        if (threadIdx.x == T) issue(q + 2 * stride, it & 1);
        // ---- X1: exchange 1
        F::template scatter<0>(x, tid, u);
        pair_slot_barrier(sub);
#pragma unroll
        for (int r = 0; r < E; r++) u[r] = xi[r * XSTR];
        RFA_PSTAMP(it, 3);
        pair_cta_barrier();
        RFA_PSTAMP(it, 4);
        // ---- C1: middle pass (its twiddles were applied before the exchange)
        Dft<16>::run(u);
        RFA_PSTAMP(it, 5);
        pair_cta_barrier();
        RFA_PSTAMP(it, 6);
        // ---- X2: exchange 2
        F::template scatter<1>(x, tid, u);
        pair_slot_barrier(sub);
#pragma unroll
        for (int r = 0; r < E; r++) u[r] = xi[r * XSTR];
        RFA_PSTAMP(it, 7);
        pair_cta_barrier();
        prev_active = active;
        prev_f = f;
    }
    if (prev_active) {
        finish(prev_f);
        publish(prev_f);
    }
    if (sub == 0) pair_cta_barrier();  // pairs with slot 1's extra barrier at the start

    // the two slots hold the same bins: reduce them in shared memory, then one atomic per bin per CTA
    if (want_peak) {
        float *red = reinterpret_cast<float *>(smem_raw);  // [SLOTS][NL], the exchange frames are free now
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; e++) red[sub * NL + F::peak_index(0, tid, e)] = worked ? pk[e] : -999999.0f;
        __syncthreads();
        for (int i = threadIdx.x; i < NL; i += G::CTA) {
            const float m = fmaxf(red[i], red[NL + i]);
            if (m > -999999.0f) atomic_max_float(p.peaks + i, m);
        }
    }
    retire_cta(p);
}
#endif  // __CUDACC__

}  // namespace rfa
