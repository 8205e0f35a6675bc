// fourstep_kernel.cuh -- the fused IQ -> spectrum path for N = 32768 and 65536 as a four-step FFT.
//
// Same reference lines as spectrum_kernel.cuh; what changes is the factorisation.  A frame of
// N = N1 * 256 points does not fit one SM's shared memory, and the residue split of spectrum_kernel
// re-reads, re-converts and re-windows the input S times.  Here, with n = 256*n1 + n2 and
// k = k1 + N1*k2:
//
//   step A (columns)  Y[k1][n2] = sum_n1 w[n] x[n] W_N1^(n1 k1)      256 FFTs of N1 points, stride 256
//                     Z[k1][n2] = Y[k1][n2] * W_N^(n2 k1)            twiddle, written row-major
//   step B (rows)     X[k1 + N1 k2] = sum_n2 Z[k1][n2] W_256^(n2 k2) N1 FFTs of 256 points
//                     dB, fft-shift, row store, peak hold
//
// Z (8 bytes per point) lives in a batch buffer of at most 128 MiB (2^24 points): most of it is still in
// the 126 MB L2 when step B reads it back; the IQ bytes are read once and the rows written once, like
// everywhere else.  The two kernels of a batch are chained by programmatic dependent launch.
// Thread layouts put neighbouring COLUMNS (step A) resp. neighbouring ROWS (step B) on neighbouring
// lanes, so that the strided side of each step is the coalesced one; per-column / per-row exchange
// buffers have an odd stride, which keeps every shared-memory access conflict-free.  A CTA keeps its
// column (row) group for the whole launch: window taps, twiddles and running peaks stay in registers.
#pragma once
#include "spectrum_kernel.cuh"
#ifdef __CUDACC__
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#endif

namespace rfa {

struct FourStepParams {
    SpectrumParams p;      // in, win, rows, row0/row_step/ring_rows/row_stride, store_from, peaks, inv_n2 (dB bias)
    const cf *tw_n1;       // per-pass Stockham twiddles of an N1-point transform (make_pass_twiddles(N1))
    const cf *tw_256;      // ... of a 256-point transform
    const cf *tw_n;        // exp(-2*pi*i*t/N), t < N
    const cf *tz;          // cluster path: [N1][256] column twiddles W_N^(n2 k1), row k1 (fourstep_cluster.cuh)
    cf *z;                 // [batch][N1][256]
    long long frame0;      // first frame of this batch
    int nbatch;            // frames in this batch
    // fused producer/consumer launch (fourstep_fused_kernel): Z is a ring of `ring` frames
    int ring;              // frames in z (two-kernel path: >= nbatch)
    int n_prod;            // CTAs 0 .. n_prod-1 transform columns, the rest rows
    unsigned int *col_done;  // [nbatch] column groups finished per frame (zeroed before the launch)
    unsigned int *row_done;  // [nbatch] row groups finished per frame
};

template <int N1>
struct GeomFS {
    static constexpr int N = N1 * 256;
    static constexpr int T1 = N1 / 16;           // threads per column transform (16 points each)
    static constexpr int CPC = 256 / T1;         // columns per CTA (step A)
    static constexpr int GROUPS_A = 256 / CPC;   // column groups
    static constexpr int GROUPS_B = N1 / 16;     // row groups of 16 rows (step B)
    static constexpr int CSTRIDE = (Plan<N1>::SMEM_POINTS | 1);   // odd: lanes = columns
    static constexpr int RSTRIDE = (Plan<256>::SMEM_POINTS | 1);  // odd: lanes = rows
    static constexpr size_t SMEM_A = (size_t)CPC * CSTRIDE * sizeof(cf);
    // staged variant: two raw tiles [N1][CPC] of b-byte IQ pairs behind the exchange buffers (128-byte aligned)
    static constexpr size_t XCHG_A = (SMEM_A + 127) / 128 * 128;
    static RFA_CX size_t tile_bytes(int bps) { return (size_t)N1 * CPC * bps; }
    static RFA_CX size_t smem_a_staged(int bps) { return XCHG_A + 2 * tile_bytes(bps); }
    // ... and, when Z leaves through a tensor-map store, one tile [N1][CPC] of finished Z points behind them
    static constexpr size_t ZTILE_BYTES = (size_t)N1 * CPC * sizeof(cf);
    static RFA_CX size_t smem_a_zstore(int bps) { return smem_a_staged(bps) + ZTILE_BYTES; }
    static constexpr size_t SMEM_B = (size_t)(2 * 16 * 256 + 16 * RSTRIDE) * sizeof(cf);  // two dense Z tiles + exchange
    static RFA_CX size_t smem_fused(int bps) { return smem_a_staged(bps) > SMEM_B ? smem_a_staged(bps) : SMEM_B; }
    static_assert(N1 == 128 || N1 == 256, "four-step covers N = 32768 and 65536");
    static_assert(Plan<N1>::PASSES == 2 && Plan<256>::PASSES == 2, "two passes per step");
};

// ---- step A: one column, thread t of its T1 threads -------------------------------------------
template <int N1, int IN>
struct FourStepA {
    using G = GeomFS<N1>;
    static constexpr int T1 = G::T1, R1 = Plan<N1>::radix(1), NB = 16 / R1;
    // raw codes of points n1 = t + r*T1 of column n2; `src` points at sample n2 + 256*t of the frame
    static RFA_HD void load_raw(const char *src, uint32_t *raw) {
#pragma unroll
        for (int r = 0; r < 16; r++)
            raw[r] = (IN == FMT_S16LE) ? ((const uint32_t *)src)[(size_t)r * T1 * 256] : (uint32_t)((const uint16_t *)src)[(size_t)r * T1 * 256];
    }
    // the same codes from a staged tile [N1 rows][CPC columns] of raw IQ pairs (the TMA box of this column group)
    static RFA_HD void load_raw_tile(const void *tile, int col, int t, uint32_t *raw) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int i = (t + r * T1) * G::CPC + col;
            raw[r] = (IN == FMT_S16LE) ? ((const uint32_t *)tile)[i] : (uint32_t)((const uint16_t *)tile)[i];
        }
    }
    static RFA_HD void load_window(const float *win, int n2, int t, float *wreg) {
#pragma unroll
        for (int r = 0; r < 16; r++) wreg[r] = (win ? win[(t + r * T1) * 256 + n2] : 1.0f) * unit_scale<IN>();
    }
    // Stockham twiddles of the second pass (frame-invariant): twreg[b*(R1-1) + r-1]
    static RFA_HD void load_pass_tw(const cf *tw_n1, int t, cf *twreg) {
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int r = 1; r < R1; r++) twreg[b * (R1 - 1) + r - 1] = tw_n1[(r - 1) * 16 + ((t + b * T1) & 15)];
    }
    // output e = b*R1 + c of this thread is k1 = (t + b*T1) + 16*c
    static RFA_HD int k1_of(int t, int e) { return (t + (e / R1) * T1) + 16 * (e % R1); }
    static RFA_HD void load_col_tw(const cf *tw_n, int n2, int t, cf *twz) {
#pragma unroll
        for (int e = 0; e < 16; e++) twz[e] = tw_n[(n2 * k1_of(t, e)) & (G::N - 1)];
    }
    static RFA_HD void first(const uint32_t *raw, const float *wreg, cf *u) {
#pragma unroll
        for (int e = 0; e < 16; e++) u[e] = decode_point<IN>(raw[e], wreg[e]);
        Dft<16>::run(u);
    }
    static RFA_HD void scatter(cf *xcol, int t, const cf *u) { pass_scatter<N1, T1, 16, 1>(xcol, t, u); }
    // second pass + column twiddle; z points at Z[frame][0][n2]
    // (`zstride` = points between consecutive k1: 256 in Z itself, CPC in the column kernel's staged store tile)
    static RFA_HD void second(const cf *xcol, const cf *twreg, const cf *twz, int t, cf *u, cf *z, int zstride = 256) {
        constexpr int STR = N1 / R1;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const cf *xi = xcol + phys(t + b * T1);
#pragma unroll
            for (int r = 0; r < R1; r++) {
                cf v = xi[r * (STR + STR / 16)];
                if (r > 0) v = cmul(v, composed_twiddle<R1>(twreg + b * (R1 - 1), r));  // R1 = 16 or 8: six (four) register twiddles per butterfly
                u[b * R1 + r] = v;
            }
            Dft<R1>::run(u + b * R1);
#pragma unroll
            for (int c = 0; c < R1; c++) {
                const int e = b * R1 + c;
                z[(size_t)k1_of(t, e) * zstride] = cmul(u[b * R1 + Dft<R1>::perm(c)], twz[e]);
            }
        }
    }
};

// ---- step B: one row k1, thread t of its 16 threads ----------------------------------------------
template <int N1>
struct FourStepB {
    using G = GeomFS<N1>;
    static RFA_HD void load_pass_tw(const cf *tw_256, int t, cf *twreg) {
#pragma unroll
        for (int r = 1; r < 16; r++) twreg[r - 1] = tw_256[(r - 1) * 16 + t];
    }
    // first pass from the staged row (points n2 = t + 16*r)
    static RFA_HD void first(const cf *zrow, int t, cf *u) {
#pragma unroll
        for (int r = 0; r < 16; r++) u[r] = zrow[t + 16 * r];
        Dft<16>::run(u);
    }
    static RFA_HD void scatter(cf *xrow, int t, const cf *u) { pass_scatter<256, 16, 16, 1>(xrow, t, u); }
    // second pass, dB, store and peak: output c is k2 = t + 16*c, bin = (k1 + N1*k2) ^ (N/2)
    template <bool PEAK, bool STORE>
    static RFA_HD void second(const cf *xrow, const cf *twreg, int t, int k1, float *out, float *pk, float db_bias) {
        cf u[16];
        const cf *xi = xrow + phys(t);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            cf v = xi[r * 17];
            if (r > 0) v = cmul(v, composed_twiddle<16>(twreg, r));
            u[r] = v;
        }
        Dft<16>::run(u);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const float db = logmag_db(u[Dft<16>::perm(c)], db_bias);
            if (STORE) out[bin_of(t, k1, c)] = db;
            if (PEAK) pk[c] = fmaxf(pk[c], db);
        }
    }
    static RFA_HD int bin_of(int t, int k1, int c) { return (k1 + N1 * (t + 16 * c)) ^ (G::N >> 1); }
};

#ifdef __CUDACC__
#ifndef RFA_FS_MINCTAS
#define RFA_FS_MINCTAS 2  // two resident column CTAs per SM (128 registers; 144 without the cap)
#endif
// 2-D tiled bulk copy (TMA): box {c0 .. , c1 ..} of the tensor map into shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint32_t bytes, unsigned long long *mbar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(mbar))
                 : "memory");
}

// grid = GROUPS_A * lanes; CTA (group, lane) transforms column group `group` of frames lane, lane+lanes, ...
// STAGED: the column group's raw IQ -- N1 rows of CPC pairs, 512 (1024) bytes apart in the frame -- arrives as ONE
// tensor-map box per frame in a two-deep ring (the per-thread version issues 16 two-byte loads per frame and
// stalls on the load/store queue: ncu lg_throttle 2.0, gpurun_out/prof_fs1).  `tmap_in` views the IQ bytes of the
// call as [frames * N1][256 * bytes-per-pair] uint8.
// 2-D tiled bulk store (TMA): a dense shared-memory tile to box {c0 .., c1 ..} of the tensor map, as one bulk group
__device__ __forceinline__ void tma_store_2d(const void *tmap, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(smem_u32(src))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// cross-CTA hand-over of the fused launch: a frame's counter is bumped once per finished group, after a CTA
// barrier and a device-scope fence (the stores of all 256 threads are then visible to whoever acquires the count)
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_count(const unsigned int *p, unsigned int want) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < want) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a lost hand-over must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ void publish_count(unsigned int *p) {  // call by ONE thread after a CTA barrier
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");  // the consumer reads Z through the async proxy (TMA)
    atomicAdd(p, 1u);
}

// One CTA of step A: column group bid % GROUPS_A of frames bid / GROUPS_A, + lanes, ...
// FUSED: Z is a ring; slot f % ring may be overwritten once frame f - ring has been read by every row group,
// and a finished group is counted in col_done[f].
// ZTMA: the finished Z points of a frame are collected in a dense tile [N1][CPC] and leave as ONE tensor-map store
// (`tmap_z` views the batch buffer as [batch * N1][512] floats) instead of 16 eight-byte stores per thread.
template <int N1, int IN, bool STAGED, bool FUSED, bool ZTMA = false>
__device__ __forceinline__ void fourstep_cols_cta(const FourStepParams &a, const CUtensorMap *tmap_in, int bid, int nblk,
                                                  unsigned char *smem_raw, unsigned long long *s_mbar,
                                                  const CUtensorMap *tmap_z = nullptr) {
    using G = GeomFS<N1>;
    using F = FourStepA<N1, IN>;
    const int col = threadIdx.x % G::CPC, t = threadIdx.x / G::CPC;
    const int group = bid % G::GROUPS_A, lane = bid / G::GROUPS_A, lanes = nblk / G::GROUPS_A;
    const int n2 = group * G::CPC + col;
    cf *xcol = reinterpret_cast<cf *>(smem_raw) + (size_t)col * G::CSTRIDE;
    constexpr int BPS = in_elem_bytes<IN>();
    constexpr uint32_t TILE_BYTES = (uint32_t)G::tile_bytes(BPS);
    unsigned char *tiles = smem_raw + G::XCHG_A;  // [2][TILE_BYTES] (STAGED)
    float wreg[16];
    cf twreg[F::NB * (F::R1 - 1)], twz[16], u[16];
    uint32_t raw[16];
    // programmatic dependent launch: the next kernel of the chain may start its (constant-table) prologue while
    // this grid drains; everything that touches Z, the IQ bytes, rows or peaks sits behind grid_dependency_wait()
    if (!FUSED) launch_dependents();
    F::load_window(a.p.win, n2, t, wreg);
    F::load_pass_tw(a.tw_n1, t, twreg);
    F::load_col_tw(a.tw_n, n2, t, twz);
    if constexpr (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(&s_mbar[0]);
            mbar_init(&s_mbar[1]);
        }
        __syncthreads();
    }
    if (!FUSED) grid_dependency_wait();
    const char *src0 = (const char *)a.p.in + ((size_t)t * 256 + n2) * BPS;
    if constexpr (STAGED) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int b = 0; b < 2; b++)
                if (lane + b * lanes < a.nbatch)
                    tma_load_2d(tiles + b * TILE_BYTES, tmap_in, group * G::CPC * BPS, (int)((a.frame0 + lane + b * lanes) * N1),
                                TILE_BYTES, &s_mbar[b]);
        }
    } else {
        if (lane < a.nbatch) F::load_raw(src0 + (a.frame0 + lane) * (long long)G::N * BPS, raw);
    }
    int it = 0;
    for (int fb = lane; fb < a.nbatch; fb += lanes, it++) {
        if (FUSED && threadIdx.x == 0 && fb >= a.ring) wait_count(a.row_done + (fb - a.ring), G::GROUPS_B);  // slot is free
        if constexpr (STAGED) {
            const int b = it & 1;
            mbar_wait(&s_mbar[b], (uint32_t)((it >> 1) & 1));
            F::load_raw_tile(tiles + b * TILE_BYTES, col, t, raw);
            F::first(raw, wreg, u);
            __syncthreads();  // the previous frame's second pass has read the exchange buffer; tile b is consumed
            if (threadIdx.x == 0 && fb + 2 * lanes < a.nbatch)
                tma_load_2d(tiles + b * TILE_BYTES, tmap_in, group * G::CPC * BPS, (int)((a.frame0 + fb + 2 * lanes) * N1), TILE_BYTES,
                            &s_mbar[b]);
        } else {
            F::first(raw, wreg, u);
            if (fb + lanes < a.nbatch) F::load_raw(src0 + (a.frame0 + fb + lanes) * (long long)G::N * BPS, raw);
            __syncthreads();  // the previous frame's second pass has read the exchange buffer
        }
        F::scatter(xcol, t, u);
        if (ZTMA && threadIdx.x == 0 && it > 0) tma_store_wait_read();  // the previous frame's store has read the tile
        __syncthreads();  // (FUSED: thread 0 passed its wait_count before the first barrier of this iteration)
        if constexpr (ZTMA) {
            cf *ztile = reinterpret_cast<cf *>(smem_raw + G::smem_a_staged(BPS));
            F::second(xcol, twreg, twz, t, u, ztile + col, G::CPC);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // these stores, then the bulk store's reads
            __syncthreads();
            if (threadIdx.x == 0) tma_store_2d(tmap_z, group * G::CPC * 2, fb * N1, ztile);
        } else {
            F::second(xcol, twreg, twz, t, u, a.z + (size_t)(FUSED ? fb % a.ring : fb) * G::N + n2);
        }
        if (FUSED) {
            __syncthreads();
            if (threadIdx.x == 0) publish_count(a.col_done + fb);
        }
    }
    if (ZTMA && threadIdx.x == 0) tma_store_wait_all();  // the last tile is in Z before the CTA leaves
}

template <int N1, int IN, bool STAGED>
__global__ void __launch_bounds__(256, RFA_FS_MINCTAS) fourstep_cols_kernel(const FourStepParams a, const __grid_constant__ CUtensorMap tmap_in) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    fourstep_cols_cta<N1, IN, STAGED, false>(a, &tmap_in, (int)blockIdx.x, (int)gridDim.x, smem_raw, s_mbar);
}

// tensor-map loads AND store
template <int N1, int IN>
__global__ void __launch_bounds__(256, RFA_FS_MINCTAS) fourstep_cols_ztma_kernel(const FourStepParams a, const __grid_constant__ CUtensorMap tmap_in,
                                                                                const __grid_constant__ CUtensorMap tmap_z) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    fourstep_cols_cta<N1, IN, true, false, true>(a, &tmap_in, (int)blockIdx.x, (int)gridDim.x, smem_raw, s_mbar, &tmap_z);
}

// grid = GROUPS_B * lanes; CTA (group, lane) transforms rows 16*group .. 16*group+15 of its frames.
// The 16 rows are 32 KB of contiguous Z: ONE bulk copy (cp.async.bulk, mbarrier completion) brings them into a
// two-deep ring of dense tiles, issued two frames ahead, so no thread ever waits on a global load of Z.  A thread
// has two identities: in the first pass lanes are consecutive POINTS of a row (the dense tile is read 128 bytes at
// a time), in the second pass lanes are consecutive ROWS (= consecutive bins k1: coalesced row stores); the
// exchange buffer between the passes is where the identity changes.
template <int N1, bool FUSED>
__device__ __forceinline__ void fourstep_rows_cta(const FourStepParams &a, int bid, int nblk, unsigned char *smem_raw,
                                                  unsigned long long *s_mbar) {
    using G = GeomFS<N1>;
    using F = FourStepB<N1>;
    const int row1 = threadIdx.x / 16, t1 = threadIdx.x % 16;  // first pass
    const int row = threadIdx.x % 16, t = threadIdx.x / 16;    // second pass
    const int group = bid % G::GROUPS_B, lane = bid / G::GROUPS_B, lanes = nblk / G::GROUPS_B;
    const int k1 = group * 16 + row;
    constexpr int TILE = 16 * 256;                          // points per tile
    cf *tile = reinterpret_cast<cf *>(smem_raw);           // [2][TILE] staged rows of Z, dense
    cf *xs = tile + 2 * TILE;                               // [16][RSTRIDE] exchange
    cf twreg[15];
    if (!FUSED) launch_dependents();
    F::load_pass_tw(a.tw_256, t, twreg);
    if (threadIdx.x == 0) {
        mbar_init(&s_mbar[0]);
        mbar_init(&s_mbar[1]);
    }
    __syncthreads();
    if (!FUSED) grid_dependency_wait();
    const cf *zg = a.z + (size_t)group * TILE;
    // FUSED: frame f sits in ring slot f % ring once all GROUPS_A column groups have counted themselves
    auto fetch = [&](int f, int b) {
        if (FUSED) {
            wait_count(a.col_done + f, G::GROUPS_A);
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores of other CTAs -> this TMA read
        }
        tma_load_1d(tile + b * TILE, zg + (size_t)(FUSED ? f % a.ring : f) * G::N, TILE * sizeof(cf), &s_mbar[b]);
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (lane + b * lanes < a.nbatch) fetch(lane + b * lanes, b);
    }
    float pk[16];
#pragma unroll
    for (int c = 0; c < 16; c++) pk[c] = -999999.0f;
    const bool want_peak = a.p.peaks != nullptr;
    bool worked = false;
    int it = 0;
    for (int fb = lane; fb < a.nbatch; fb += lanes, it++) {
        const long long f = a.frame0 + fb;
        const int b = it & 1;
        mbar_wait(&s_mbar[b], (uint32_t)((it >> 1) & 1));
        cf u[16];
        F::first(tile + b * TILE + row1 * 256, t1, u);
        __syncthreads();  // the previous frame's second pass is done with xs; everybody has read tile b
        if (threadIdx.x == 0) {
            if (FUSED) {  // this row group no longer needs the frame's ring slot
                __threadfence();
                atomicAdd(a.row_done + fb, 1u);
            }
            if (fb + 2 * lanes < a.nbatch) fetch(fb + 2 * lanes, b);
        }
        F::scatter(xs + row1 * G::RSTRIDE, t1, u);
        __syncthreads();
        float *out = a.p.rows + frame_row(a.p, f) * a.p.row_stride;
        const cf *xrow = xs + row * G::RSTRIDE;
        if (f >= a.p.store_from) {
            if (want_peak)
                F::template second<true, true>(xrow, twreg, t, k1, out, pk, a.p.inv_n2);
            else
                F::template second<false, true>(xrow, twreg, t, k1, out, pk, a.p.inv_n2);
        } else if (want_peak) {
            F::template second<true, false>(xrow, twreg, t, k1, out, pk, a.p.inv_n2);
        }
        worked = true;
    }
    if (want_peak && worked) {
#pragma unroll
        for (int c = 0; c < 16; c++) atomic_max_float(a.p.peaks + F::bin_of(t, k1, c), pk[c]);
    }
}

template <int N1>
__global__ void __launch_bounds__(256) fourstep_rows_kernel(const FourStepParams a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    fourstep_rows_cta<N1, false>(a, (int)blockIdx.x, (int)gridDim.x, smem_raw, s_mbar);
}

#ifdef RFA_LAB  // lab builds only: slower than the two-kernel path, and its hand-over watchdog traps the context
// Both steps in ONE cooperative launch (all CTAs co-resident): CTAs 0 .. n_prod-1 produce column transforms into
// a ring of Z frames that fits the L2, the others consume them as row transforms, frame by frame, handing over
// through per-frame counters.  Z never travels to HBM, the two steps overlap on every SM (one is store-, the
// other load-heavy), and there is one prologue and one tail per call instead of one per batch and step.
template <int N1, int IN>
__global__ void __launch_bounds__(256, 2) fourstep_fused_kernel(const FourStepParams a, const __grid_constant__ CUtensorMap tmap_in) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    if ((int)blockIdx.x < a.n_prod)
        fourstep_cols_cta<N1, IN, true, true>(a, &tmap_in, (int)blockIdx.x, a.n_prod, smem_raw, s_mbar);
    else
        fourstep_rows_cta<N1, true>(a, (int)blockIdx.x - a.n_prod, (int)gridDim.x - a.n_prod, smem_raw, s_mbar);
}
#endif  // RFA_LAB
#endif  // __CUDACC__

}  // namespace rfa
