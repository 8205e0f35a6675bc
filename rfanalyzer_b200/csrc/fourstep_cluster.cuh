// fourstep_cluster.cuh -- N = 32768 / 65536 on thread-block clusters: the four-step transform of fourstep_kernel.cuh
// with its intermediate Z in DISTRIBUTED SHARED MEMORY instead of HBM.
//
// Same reference lines as spectrum_kernel.cuh (the whole frame in one call: pffft.c:1904 pffft_transform_ordered,
// nativedsp.cpp:44-81).  A frame of N = N1 * 256 points (256 or 512 KB of complex floats) does not fit one SM, so a
// cluster of CS = N1 / 64 CTAs (4 for 65536, 2 for 32768) holds it: with n = 256*n1 + n2, k = k1 + N1*k2 as before,
//
//   step A   CTA `rank` owns columns n2 = rank*256/CS ...: convert, window, N1-point column transforms, multiply by
//            W_N^(n2 k1), and every finished point Z[k1][n2] is written straight into the shared memory of the CTA
//            that owns row k1 (st.shared::cluster, 128 contiguous bytes per half-warp);
//   -- one cluster barrier (arrive.release / wait.acquire) --
//   step B   CTA `rank` owns rows k1 = 64*rank .. 64*rank+63: 256-point row transforms out of its own shared memory,
//            dB, fft-shift, row store, peak hold.
//
// An IQ sample is read from HBM once (2-D tensor-map boxes, one per column group and frame) and one float per bin is
// written; nothing else touches global memory except the window taps and the column twiddles W_N^(n2 k1), which are
// read per frame from tables that live in L2.  The two-kernel path moves 16 more bytes per sample through L2 / HBM.
//
// A CTA is two teams of 256 threads; each team transforms two of the CTA's four column groups and then two of its
// four row groups per frame, synchronising with named barriers of its own, so the teams drift apart and overlap each
// other's exchange and butterfly phases.  The second cluster barrier of a frame ("everybody has read Z, the next
// frame may overwrite it") is split: arrive after a thread's last read of Z, wait just before its first remote store
// of the next frame, one and a half unit transforms later.
//
// Shared memory per CTA (229376 bytes): Z tile [64 rows][256] (128 KB), one exchange buffer of 4096 points per team
// (dense, no padding: step A keeps neighbouring COLUMNS on neighbouring lanes, step B XOR-swizzles the row slot),
// raw IQ tiles (32 KB).
#pragma once
#include "fourstep_kernel.cuh"

namespace rfa {

template <int N1, int IN>
struct ClusterFS {
    using G = GeomFS<N1>;
    using FA = FourStepA<N1, IN>;
    static constexpr int N = G::N, T1 = G::T1, CPC = G::CPC, R1 = FA::R1, NB = FA::NB;
    static constexpr int CS = N1 / 64;      // CTAs per cluster
    static constexpr int COLS = 256 / CS;   // columns per CTA (step A)
    static constexpr int ROWS = 64;         // rows k1 per CTA (step B)
    static constexpr int UNITS = 4;         // column groups per CTA = row groups per CTA
    static constexpr int BPS = in_elem_bytes<IN>();
    static constexpr int TILE_BYTES = N1 * CPC * BPS;  // raw IQ of one column group of one frame
    static constexpr int NT = BPS == 2 ? 2 : 1;        // tile buffers per team
    static constexpr size_t Z_BYTES = (size_t)ROWS * 256 * sizeof(cf);
    static constexpr size_t XCH_BYTES = 4096 * sizeof(cf);
    static constexpr size_t OFF_XCH = Z_BYTES, OFF_RAW = OFF_XCH + 2 * XCH_BYTES;
    static constexpr size_t SMEM = OFF_RAW + 2 * (size_t)NT * TILE_BYTES;
    static_assert(COLS / CPC == UNITS && ROWS / 16 == UNITS, "four column groups and four row groups per CTA");
    static_assert(N1 * CPC == 4096, "a unit is 4096 points = 256 threads x 16");

    // ---- step A, one column group: thread (col, t) ----
    // exchange layout: logical point p of column col at xa[p * CPC + col] (a half-warp = 16 neighbouring columns)
    static RFA_HD void a_scatter(cf *xa, int col, int t, const cf *u) {
#pragma unroll
        for (int c = 0; c < 16; c++) xa[(16 * t + c) * CPC + col] = u[Dft<16>::perm(c)];
    }
    // column twiddles of this thread's 16 outputs from the table tz[k1][n2] = W_N^(n2 k1)
    static RFA_HD void load_col_tw(const cf *tz, int n2, int t, cf *twz) {
#pragma unroll
        for (int e = 0; e < 16; e++) twz[e] = tz[(size_t)FA::k1_of(t, e) * 256 + n2];
    }
    // second pass + column twiddle; `tw1` = pass-1 table of the N1-point plan, put(owner CTA, local row, value)
    template <class Put>
    static RFA_HD void a_second(const cf *xa, const cf *tw1, const cf *twz, int col, int t, cf *u, Put put) {
        constexpr int STR = N1 / R1;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int i = t + b * T1;  // < 16
            const cf *xi = xa + i * CPC + col;
            cf t[R1 - 1];  // the entries composed_twiddle reads: r = 1, 2, 3 and 4 (8, 12) -- the same products as the two-kernel path
#pragma unroll
            for (int r = 1; r < R1; r++)
                if (r < 4 || (r & 3) == 0) t[r - 1] = tw1[(r - 1) * 16 + (i & 15)];
#pragma unroll
            for (int r = 0; r < R1; r++) {
                cf v = xi[r * STR * CPC];
                if (r > 0) v = cmul(v, composed_twiddle<R1>(t, r));
                u[b * R1 + r] = v;
            }
            Dft<R1>::run(u + b * R1);
        }
#pragma unroll
        for (int c = 0; c < R1; c++) {  // k1 = i + 16*c: owner c / 4, local row i + 16*(c % 4)
#pragma unroll
            for (int b = 0; b < NB; b++)
                put(c >> 2, t + b * T1 + 16 * (c & 3), cmul(u[b * R1 + Dft<R1>::perm(c)], twz[b * R1 + c]));
        }
    }

    // ---- step B, one group of 16 rows ----
    // first pass: thread (row1, t1) takes points t1 + 16 r of its row from the CTA's Z tile
    static RFA_HD void b_first(const cf *zrow, int t1, cf *u) {
#pragma unroll
        for (int r = 0; r < 16; r++) u[r] = zrow[t1 + 16 * r];
        Dft<16>::run(u);
    }
    // exchange layout: logical point p of row q at xb[p*16 + (q ^ (p >> 4))] -- the writers of a half-warp differ
    // in p >> 4, the readers in q: both sides touch 16 distinct slots of 8 bytes
    static RFA_HD void b_scatter(cf *xb, int row1, int t1, const cf *u) {
        cf *y = xb + 256 * t1 + (row1 ^ t1);
#pragma unroll
        for (int c = 0; c < 16; c++) y[16 * c] = u[Dft<16>::perm(c)];
    }
    static RFA_HD int bin_of(int t, int k1, int c) { return (k1 + N1 * (t + 16 * c)) ^ (N >> 1); }
    // second pass, dB, store, peak: thread (row, t), `tw256` = pass-1 table of the 256-point plan
    template <bool PEAK, bool STORE>
    static RFA_HD void b_second(const cf *xb, const cf *tw256, int row, int t, int k1, float *out, float *pk, float db_bias) {
        cf u[16];
        const cf *xi = xb + 16 * t;
        cf tw[15];
#pragma unroll
        for (int r = 1; r < 16; r++)
            if (r < 4 || (r & 3) == 0) tw[r - 1] = tw256[(r - 1) * 16 + t];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            cf v = xi[256 * r + (row ^ r)];
            if (r > 0) v = cmul(v, composed_twiddle<16>(tw, r));
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
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t addr, uint32_t rank) {  // mapa: the same offset in a peer CTA
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, cf v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
#ifdef RFA_CL_NOBAR
__device__ __forceinline__ void cluster_arrive() {}
__device__ __forceinline__ void cluster_wait() {}
#else
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#endif
__device__ __forceinline__ void team_barrier(int team) { asm volatile("bar.sync %0, 256;" ::"r"(team + 1) : "memory"); }

// grid = clusters * CS CTAs of 512 threads, cluster dimension CS; cluster q transforms frames q, q + clusters, ...
template <int N1, int IN>
__global__ void __launch_bounds__(512, 1) fourstep_cluster_kernel(const FourStepParams a, const __grid_constant__ CUtensorMap tmap_in) {
    using C = ClusterFS<N1, IN>;
    using FA = FourStepA<N1, IN>;
    constexpr int CS = C::CS, CPC = C::CPC, NT = C::NT, BPS = C::BPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2][2];
    __shared__ cf s_tw256[15 * 16];
    __shared__ cf s_tw1[N1 == 256 ? 1 : (FA::R1 - 1) * 16];
    const int team = threadIdx.x >> 8, tid = threadIdx.x & 255;
    const int rank = (int)cluster_ctarank();
    const int cluster = (int)blockIdx.x / CS, clusters = (int)gridDim.x / CS;
    cf *zloc = reinterpret_cast<cf *>(smem_raw);
    cf *xch = reinterpret_cast<cf *>(smem_raw + C::OFF_XCH) + team * 4096;
    unsigned char *tiles = smem_raw + C::OFF_RAW + (size_t)team * NT * C::TILE_BYTES;
    for (int i = threadIdx.x; i < 15 * 16; i += 512) s_tw256[i] = a.tw_256[i];
    if (N1 != 256)
        for (int i = threadIdx.x; i < (FA::R1 - 1) * 16; i += 512) s_tw1[i] = a.tw_n1[i];
    const cf *tw1 = N1 == 256 ? s_tw256 : s_tw1;  // a 256-point column plan has the row plan's table
    const int col = tid % CPC, t = tid / CPC;     // step A
    const int row1 = tid / 16, t1 = tid % 16;     // step B, first pass: lanes = points
    const int row = tid % 16, tb = tid / 16;      // step B, second pass: lanes = rows
    uint32_t zpeer[CS];
#pragma unroll
    for (int r = 0; r < CS; r++) zpeer[r] = map_shared_rank(smem_u32(zloc), (uint32_t)r);
    if (tid == 0)
        for (int b = 0; b < NT; b++) mbar_init(&s_mbar[team][b]);
    __syncthreads();
    // tile `seq` of this team: unit seq & 1 of the team's frame seq >> 1
    auto fetch = [&](int seq) {
        const long long fb = cluster + (long long)(seq >> 1) * clusters;
        if (fb < a.nbatch) {
            const int g = 2 * (seq & 1) + team;
            tma_load_2d(tiles + (seq % NT) * C::TILE_BYTES, &tmap_in, (rank * C::COLS + g * CPC) * BPS, (int)((a.frame0 + fb) * N1),
                        (uint32_t)C::TILE_BYTES, &s_mbar[team][seq % NT]);
        }
    };
    if (tid == 0)
        for (int b = 0; b < NT; b++) fetch(b);
    // nobody stores into a peer's shared memory before that CTA runs
    cluster_arrive();
    cluster_wait();
    float pk[2][16];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int c = 0; c < 16; c++) pk[j][c] = -999999.0f;
    const bool want_peak = a.p.peaks != nullptr;
    float wreg[16];  // window taps (times the format's unit) of the unit about to start
#ifdef RFA_CL_NOWIN
#pragma unroll
    for (int r = 0; r < 16; r++) wreg[r] = 0.001f * (float)col;
#else
    FA::load_window(a.p.win, rank * C::COLS + team * CPC + col, t, wreg);
#endif
    int it = 0;
    for (long long fb = cluster; fb < a.nbatch; fb += clusters, it++) {
        // ---- step A: two column groups per team ----
#pragma unroll 1
        for (int j = 0; j < 2; j++) {
            const int g = 2 * j + team, n2 = rank * C::COLS + g * CPC + col, seq = 2 * it + j;
            cf u[16];
            {
                uint32_t raw[16];
                mbar_wait(&s_mbar[team][seq % NT], (uint32_t)((seq / NT) & 1));
                FA::load_raw_tile(tiles + (seq % NT) * C::TILE_BYTES, col, t, raw);
                FA::first(raw, wreg, u);
            }
            // the column twiddles travel from L2 while the exchange runs
            cf twz[16];
#ifdef RFA_CL_NOTWZ
#pragma unroll
            for (int e = 0; e < 16; e++) twz[e] = cf{0.5f, 0.001f * (float)n2};
#else
            C::load_col_tw(a.tz, n2, t, twz);
#endif
            team_barrier(team);  // the tile is consumed, the previous unit is done with the exchange buffer
            if (tid == 0) fetch(seq + NT);
            C::a_scatter(xch, col, t, u);
            team_barrier(team);
            if (j == 0 && it > 0) cluster_wait();  // every CTA has read the previous frame out of its Z tile
            const uint32_t zoff = (uint32_t)((t * 256 + n2) * sizeof(cf));
            auto put = [&](int owner, int lrow, cf v) {
                // lrow = i + 16*(c % 4) with i = t + b*T1: the part that depends on t sits in zoff
#ifdef RFA_CL_LOCAL
                zloc[(size_t)lrow * 256 + n2] = v;
#else
                st_cluster(zpeer[owner] + zoff + (uint32_t)((lrow - t) * 256 * sizeof(cf)), v);
#endif
            };
            C::a_second(xch, tw1, twz, col, t, u, put);
            // window taps of the team's NEXT column group (the other one): in flight across the rest of this unit's tail,
            // and, after the frame's second unit, across step B
#ifndef RFA_CL_NOWIN
            FA::load_window(a.p.win, rank * C::COLS + (2 * (1 - j) + team) * CPC + col, t, wreg);
#endif
        }
        cluster_arrive();  // my points of Z are written ...
        cluster_wait();    // ... and so are everybody else's
        // ---- step B: two row groups per team ----
        const long long f = a.frame0 + fb;
        float *out = a.p.rows + frame_row(a.p, f) * a.p.row_stride;
        const bool store = f >= a.p.store_from;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int g = 2 * j + team;
            cf u[16];
            C::b_first(zloc + (size_t)(16 * g + row1) * 256, t1, u);
            if (j == 1) cluster_arrive();  // this thread no longer needs the frame's Z
            team_barrier(team);
            C::b_scatter(xch, row1, t1, u);
            team_barrier(team);
            const int k1 = rank * C::ROWS + 16 * g + row;
            if (store) {
                if (want_peak)
                    C::template b_second<true, true>(xch, s_tw256, row, tb, k1, out, pk[j], a.p.inv_n2);
                else
                    C::template b_second<false, true>(xch, s_tw256, row, tb, k1, out, pk[j], a.p.inv_n2);
            } else if (want_peak) {
                C::template b_second<true, false>(xch, s_tw256, row, tb, k1, out, pk[j], a.p.inv_n2);
            }
        }
    }
    if (it > 0) {
        cluster_wait();  // pairs with the last frame's arrive; no CTA leaves while a peer may still address it
        if (want_peak) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int k1 = rank * C::ROWS + 16 * (2 * j + team) + row;
#pragma unroll
                for (int c = 0; c < 16; c++) atomic_max_float(a.p.peaks + C::bin_of(tb, k1, c), pk[j][c]);
            }
        }
    }
}
#endif  // __CUDACC__

}  // namespace rfa
