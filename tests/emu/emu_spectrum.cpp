// emu_spectrum.cpp -- CPU emulation of the fused spectrum kernel (test support).
//
// There is no GPU in the build container, so the kernel's per-thread phase functions
// (rfanalyzer_b200/csrc/spectrum_kernel.cuh, all __host__ __device__) are run here thread
// by thread, with loop boundaries standing in for __syncthreads().  This exercises the exact
// index arithmetic, twiddle addressing, butterflies, conversion and dB scaling the GPU runs.
// It is NOT a product path: nothing in rfanalyzer_b200 loads it.
#include <array>
#include <cstring>
#include <vector>

#include "../../rfanalyzer_b200/csrc/rfa_tables.h"
#include "../../rfanalyzer_b200/csrc/spectrum_kernel.cuh"
#include "../../rfanalyzer_b200/csrc/spectrum2_kernel.cuh"
#include "../../rfanalyzer_b200/csrc/spectrum64_kernel.cuh"
#include "../../rfanalyzer_b200/csrc/fourstep_kernel.cuh"
#include "../../rfanalyzer_b200/csrc/fourstep_cluster.cuh"

using namespace rfa;

namespace {

template <int NL, int S, int IN, int OUT, int PASS>
struct EmuPasses {
    using F = SpectrumFrame<NL, S, IN, OUT>;
    using TwRegs = std::vector<std::array<cf, (F::LAST_TW_REG ? F::LAST_TW : 1)>>;
    static void run(std::vector<cf> &x, const cf *tw, TwRegs &TWR, std::vector<std::array<cf, Geom<NL>::E>> &U) {
        constexpr int T = Geom<NL>::T;
        if constexpr (PASS < Plan<NL>::PASSES) {
            for (int tid = 0; tid < T; tid++) F::template scatter<PASS - 1>(x.data(), tid, U[tid].data());
            for (int tid = 0; tid < T; tid++) {
                if constexpr (PASS == F::LAST && F::LAST_TW_REG)
                    F::gather_last_reg(x.data(), TWR[tid].data(), tid, U[tid].data());
                else
                    F::template gather<PASS>(x.data(), tw, tid, U[tid].data());
            }
            EmuPasses<NL, S, IN, OUT, PASS + 1>::run(x, tw, TWR, U);
        }
    }
};

template <int NL, int S, int IN, int OUT>
int emu_one(SpectrumParams p, float *peaks) {
    using G = Geom<NL>;
    using F = SpectrumFrame<NL, S, IN, OUT>;
    constexpr int T = G::T, E = G::E, N = NL * S;
    std::vector<cf> tw = make_pass_twiddles(NL), twN(N);
    make_twiddles(N, twN.data());
    p.tw = tw.data();
    p.twN = twN.data();
    p.inv_n2 = -3.0102999566398120f * log2f((float)N);  // dB bias, see logmag_db
    std::vector<cf> x(Plan<NL>::SMEM_POINTS);
    std::vector<std::array<cf, E>> U(T);
    std::vector<std::array<float, E>> PK((size_t)T * S);
    for (auto &a : PK) a.fill(-999999.0f);
    std::vector<std::array<float, E>> W(T);
    if (S == 1) {
        constexpr int R = Plan<NL>::radix(0), STR = NL / R;
        for (int tid = 0; tid < T; tid++)
            for (int b = 0; b < E / R; b++)
                for (int r = 0; r < R; r++)
                    W[tid][b * R + r] = (p.win ? p.win[tid + b * T + r * STR] : 1.0f) * unit_scale<IN>();
    }
    std::vector<std::array<cf, (F::LAST_TW_REG ? F::LAST_TW : 1)>> TWR(T);
    if constexpr (F::LAST_TW_REG)
        for (int tid = 0; tid < T; tid++) F::load_last_tw(tw.data(), tid, TWR[tid].data());
    std::vector<std::array<uint32_t, E>> RAW(T);
    for (long long f = 0; f < p.nframes; f++)
        for (int c = 0; c < S; c++) {
            for (int tid = 0; tid < T; tid++) {
                if constexpr (F::PREFETCH) {
                    F::load_raw((const char *)p.in + (f * (long long)N + tid) * in_elem_bytes<IN>(), RAW[tid].data());
                    F::first_from_raw(RAW[tid].data(), W[tid].data(), U[tid].data());
                } else {
                    F::first(p, f, c, tid, W[tid].data(), U[tid].data());
                }
            }
            if constexpr (Plan<NL>::PASSES > 1) EmuPasses<NL, S, IN, OUT, 1>::run(x, tw.data(), TWR, U);
            for (int tid = 0; tid < T; tid++) {
                float *out = p.rows + frame_row(p, f) * p.row_stride;
                if (OUT == OUT_DB && peaks)
                    F::template emit<true, true>(out, c, tid, U[tid].data(), PK[(size_t)c * T + tid].data(), p.inv_n2);
                else
                    F::template emit<false, true>(out, c, tid, U[tid].data(), nullptr, p.inv_n2);
            }
        }
    if (OUT == OUT_DB && peaks) {
        for (int i = 0; i < N; i++) peaks[i] = -999999.0f;
        for (int c = 0; c < S; c++)
            for (int tid = 0; tid < T; tid++)
                for (int e = 0; e < E; e++) {
                    float &dst = peaks[F::peak_index(c, tid, e)];
                    dst = dst > PK[(size_t)c * T + tid][e] ? dst : PK[(size_t)c * T + tid][e];
                }
    }
    if (OUT == OUT_DB && p.avg)
        for (int i = 0; i < N; i++) p.avg[i] = boxcar_average(p, i, [](const float *a) { return *a; });
    return 0;
}

template <int NL, int S>
int emu_size(int in_fmt, int out_kind, const SpectrumParams &p, float *peaks) {
    if (out_kind == OUT_CPLX) return in_fmt == FMT_CF32 ? emu_one<NL, S, FMT_CF32, OUT_CPLX>(p, peaks) : -1;
    switch (in_fmt) {
        case FMT_S8: return emu_one<NL, S, FMT_S8, OUT_DB>(p, peaks);
        case FMT_U8: return emu_one<NL, S, FMT_U8, OUT_DB>(p, peaks);
        case FMT_S16LE: return emu_one<NL, S, FMT_S16LE, OUT_DB>(p, peaks);
        case FMT_CF32: return emu_one<NL, S, FMT_CF32, OUT_DB>(p, peaks);
        case FMT_PF32: return emu_one<NL, S, FMT_PF32, OUT_DB>(p, peaks);
    }
    return -1;
}

// ---- dual-frame kernel (spectrum2_kernel.cuh): same schedule of frame pairs, thread by thread ----
template <int NL, int IN, int PASS>
struct EmuPasses2 {
    using F = SpectrumFrame2<NL, IN>;
    static void run(std::vector<c2> &x, const cf *tw, std::vector<std::array<cf, F::LAST_TW>> &TWR,
                    std::vector<std::array<c2, 16>> &U) {
        constexpr int T = Geom2<NL>::T;
        if constexpr (PASS < Plan<NL>::PASSES) {
            for (int tid = 0; tid < T; tid++) F::template scatter<PASS - 1>(x.data(), tid, U[tid].data());
            for (int tid = 0; tid < T; tid++) {
                if constexpr (PASS == F::LAST)
                    F::gather_last_reg(x.data(), TWR[tid].data(), tid, U[tid].data());
                else
                    F::template gather<PASS>(x.data(), tw, tid, U[tid].data());
            }
            EmuPasses2<NL, IN, PASS + 1>::run(x, tw, TWR, U);
        }
    }
};

template <int NL, int IN>
int emu_two(SpectrumParams p, float *peaks) {
    using G = Geom2<NL>;
    using F = SpectrumFrame2<NL, IN>;
    constexpr int T = G::T, E = G::E, N = NL;
    std::vector<cf> tw = make_pass_twiddles(NL);
    p.tw = tw.data();
    p.inv_n2 = -3.0102999566398120f * log2f((float)N);
    std::vector<c2> x(Plan<NL>::SMEM_POINTS);
    std::vector<std::array<c2, 16>> U(T);
    std::vector<std::array<float, 16>> PK(T), W(T);
    for (auto &a : PK) a.fill(-999999.0f);
    for (int tid = 0; tid < T; tid++)
        for (int r = 0; r < E; r++) W[tid][r] = (p.win ? p.win[tid + r * (NL / 16)] : 1.0f) * unit_scale<IN>();
    std::vector<std::array<cf, F::LAST_TW>> TWR(T);
    for (int tid = 0; tid < T; tid++) F::load_last_tw(tw.data(), tid, TWR[tid].data());
    const long long npairs = (p.nframes + 1) / 2;
    constexpr int BPS = in_elem_bytes<IN>();
    for (long long v = 0; v < npairs; v++) {
        const long long fB = p.nframes - 1 - 2 * v, fA = fB - 1;
        for (int tid = 0; tid < T; tid++) {
            uint32_t rawA[16], rawB[16];
            F::load_raw((const char *)p.in + (fB * (long long)N + tid) * BPS, rawB);
            F::load_raw((const char *)p.in + ((fA >= 0 ? fA : fB) * (long long)N + tid) * BPS, rawA);
            F::first_from_raw(rawA, rawB, W[tid].data(), U[tid].data());
        }
        EmuPasses2<NL, IN, 1>::run(x, tw.data(), TWR, U);
        for (int tid = 0; tid < T; tid++) {
            float *outB = p.rows + frame_row(p, fB) * p.row_stride;
            float *outA = p.rows + frame_row(p, fA >= 0 ? fA : fB) * p.row_stride;
            const bool storeB = fB >= p.store_from, storeA = fA >= 0 && fA >= p.store_from;
            if (peaks)
                F::template emit<true>(outA, outB, storeA, storeB, tid, U[tid].data(), PK[tid].data(), p.inv_n2);
            else
                F::template emit<false>(outA, outB, storeA, storeB, tid, U[tid].data(), nullptr, p.inv_n2);
        }
    }
    if (peaks) {
        for (int i = 0; i < N; i++) peaks[i] = -999999.0f;
        for (int tid = 0; tid < T; tid++)
            for (int e = 0; e < E; e++) {
                float &dst = peaks[F::peak_index(tid, e)];
                dst = dst > PK[tid][e] ? dst : PK[tid][e];
            }
    }
    if (p.avg)
        for (int i = 0; i < N; i++) p.avg[i] = boxcar_average(p, i, [](const float *a) { return *a; });
    return 0;
}

template <int NL>
int emu_two_fmt(int in_fmt, const SpectrumParams &p, float *peaks) {
    switch (in_fmt) {
        case FMT_S8: return emu_two<NL, FMT_S8>(p, peaks);
        case FMT_U8: return emu_two<NL, FMT_U8>(p, peaks);
        case FMT_S16LE: return emu_two<NL, FMT_S16LE>(p, peaks);
    }
    return -1;
}

}  // namespace

static int emu_dispatch(int N, int in_fmt, int out_kind, int window_kind, const void *in, const float *in_im,
                        long long nframes, float *rows, float *peaks, float *avg, int L, bool dual, long long store_from);

extern "C" int emu_spectrum_avg(int N, int in_fmt, int out_kind, int window_kind /* -1: none */, const void *in,
                                const float *in_im, long long nframes, float *rows, float *peaks, float *avg, int L) {
    return emu_dispatch(N, in_fmt, out_kind, window_kind, in, in_im, nframes, rows, peaks, avg, L, false, 0);
}

// ---- two-pass 64 x 64 kernel (spectrum64_kernel.cuh), N = 4096 ----
template <int IN>
int emu_k64(const void *in, int window_kind, long long nframes, float *rows, float *peaks) {
    using F = SpectrumFrame64<IN>;
    constexpr int N = 4096, ROW = Geom64::ROW;
    std::vector<cf> twN(N), tw(64 * 64), x(64 * ROW);
    make_twiddles(N, twN.data());
    for (int i = 0; i < 64 * 64; i++) tw[i] = twN[((i >> 6) * (i & 63)) & 4095];
    std::vector<float> w(N, unit_scale<IN>());
    if (window_kind >= 0) {
        make_window(window_kind, N, w.data());
        for (auto &v : w) v *= unit_scale<IN>();
    }
    std::vector<std::array<float, 64>> PK(64);
    for (auto &a : PK) a.fill(-999999.0f);
    const float bias = -3.0102999566398120f * log2f((float)N);
    for (long long f = 0; f < nframes; f++) {
        for (int t = 0; t < 64; t++)
            F::pass_a((const char *)in + (f * (long long)N + t) * in_elem_bytes<IN>(), w.data() + t, tw.data() + t, x.data() + t);
        for (int t = 0; t < 64; t++) F::template pass_b<true, true>(x.data() + t * ROW, rows + f * N + t, PK[t].data(), bias);
    }
    for (int t = 0; t < 64; t++)
        for (int c = 0; c < 64; c++) peaks[F::peak_index(t, c)] = PK[t][c];
    return 0;
}

extern "C" int emu_spectrum64(int in_fmt, int window_kind, const void *in, long long nframes, float *rows, float *peaks) {
    switch (in_fmt) {
        case FMT_S8: return emu_k64<FMT_S8>(in, window_kind, nframes, rows, peaks);
        case FMT_U8: return emu_k64<FMT_U8>(in, window_kind, nframes, rows, peaks);
        case FMT_S16LE: return emu_k64<FMT_S16LE>(in, window_kind, nframes, rows, peaks);
    }
    return -1;
}

// ---- four-step path (fourstep_kernel.cuh), N = 32768 / 65536: columns -> Z -> rows, thread by thread ----
template <int N1, int IN>
int emu_fourstep(const void *in, int window_kind, long long nframes, float *rows, float *peaks, long long store_from) {
    using G = GeomFS<N1>;
    using FA = FourStepA<N1, IN>;
    using FB = FourStepB<N1>;
    constexpr int N = G::N, T1 = G::T1, BPS = in_elem_bytes<IN>();
    std::vector<cf> tw1 = make_pass_twiddles(N1), tw256 = make_pass_twiddles(256), twN(N);
    make_twiddles(N, twN.data());
    std::vector<float> win;
    if (window_kind >= 0) {
        win.resize(N);
        make_window(window_kind, N, win.data());
    }
    const float bias = -3.0102999566398120f * log2f((float)N);
    std::vector<cf> z((size_t)N), xcol(G::CSTRIDE), xrow(G::RSTRIDE);
    std::vector<std::array<cf, 16>> U(T1 > 16 ? T1 : 16);
    std::vector<std::array<float, 16>> PK((size_t)N1 * 16);  // [k1][t]
    for (auto &a : PK) a.fill(-999999.0f);
    for (long long f = 0; f < nframes; f++) {
        const char *frame = (const char *)in + f * (long long)N * BPS;
        for (int n2 = 0; n2 < 256; n2++) {
            for (int t = 0; t < T1; t++) {
                float wreg[16];
                uint32_t raw[16];
                FA::load_window(win.empty() ? nullptr : win.data(), n2, t, wreg);
                FA::load_raw(frame + ((size_t)t * 256 + n2) * BPS, raw);
                FA::first(raw, wreg, U[t].data());
                FA::scatter(xcol.data(), t, U[t].data());
            }
            for (int t = 0; t < T1; t++) {
                cf twreg[FA::NB * (FA::R1 - 1)], twz[16], u[16];
                FA::load_pass_tw(tw1.data(), t, twreg);
                FA::load_col_tw(twN.data(), n2, t, twz);
                FA::second(xcol.data(), twreg, twz, t, u, z.data() + n2);
            }
        }
        float *out = rows + f * (long long)N;
        for (int k1 = 0; k1 < N1; k1++) {
            for (int t = 0; t < 16; t++) {
                FB::first(z.data() + (size_t)k1 * 256, t, U[t].data());
                FB::scatter(xrow.data(), t, U[t].data());
            }
            for (int t = 0; t < 16; t++) {
                cf twreg[15];
                FB::load_pass_tw(tw256.data(), t, twreg);
                float *pk = PK[(size_t)k1 * 16 + t].data();
                if (f >= store_from)
                    FB::template second<true, true>(xrow.data(), twreg, t, k1, out, pk, bias);
                else
                    FB::template second<true, false>(xrow.data(), twreg, t, k1, out, pk, bias);
            }
        }
    }
    if (peaks)
        for (int k1 = 0; k1 < N1; k1++)
            for (int t = 0; t < 16; t++)
                for (int c = 0; c < 16; c++) peaks[FB::bin_of(t, k1, c)] = PK[(size_t)k1 * 16 + t][c];
    return 0;
}

template <int N1>
int emu_fourstep_fmt(int in_fmt, const void *in, int window_kind, long long nframes, float *rows, float *peaks, long long store_from) {
    switch (in_fmt) {
        case FMT_S8: return emu_fourstep<N1, FMT_S8>(in, window_kind, nframes, rows, peaks, store_from);
        case FMT_U8: return emu_fourstep<N1, FMT_U8>(in, window_kind, nframes, rows, peaks, store_from);
        case FMT_S16LE: return emu_fourstep<N1, FMT_S16LE>(in, window_kind, nframes, rows, peaks, store_from);
    }
    return -1;
}

extern "C" int emu_fourstep_spectrum(int N, int in_fmt, int window_kind, const void *in, long long nframes, float *rows,
                                     float *peaks, long long store_from) {
    if (N == 65536) return emu_fourstep_fmt<256>(in_fmt, in, window_kind, nframes, rows, peaks, store_from);
    if (N == 32768) return emu_fourstep_fmt<128>(in_fmt, in, window_kind, nframes, rows, peaks, store_from);
    return -1;
}

// ---- cluster path (fourstep_cluster.cuh): CS CTAs, each with its own Z tile; thread by thread, unit by unit ----
template <int N1, int IN>
int emu_cluster(const void *in, int window_kind, long long nframes, float *rows, float *peaks, long long store_from) {
    using C = ClusterFS<N1, IN>;
    using FA = FourStepA<N1, IN>;
    constexpr int N = C::N, CS = C::CS, CPC = C::CPC, BPS = C::BPS;
    std::vector<cf> tw1 = make_pass_twiddles(N1), tw256 = make_pass_twiddles(256), twN(N), tz((size_t)N);
    make_twiddles(N, twN.data());
    for (int k1 = 0; k1 < N1; k1++)
        for (int n2 = 0; n2 < 256; n2++) tz[(size_t)k1 * 256 + n2] = twN[(size_t)(n2 * k1) & (size_t)(N - 1)];
    std::vector<float> win;
    if (window_kind >= 0) {
        win.resize(N);
        make_window(window_kind, N, win.data());
    }
    const float bias = -3.0102999566398120f * log2f((float)N);
    std::vector<std::vector<cf>> Z(CS, std::vector<cf>((size_t)C::ROWS * 256));
    std::vector<cf> xch(4096);
    std::vector<unsigned char> tile((size_t)C::TILE_BYTES);
    std::vector<std::array<cf, 16>> U(256);
    std::vector<std::array<float, 16>> PK((size_t)CS * C::UNITS * 256);  // [rank][g][tid]
    for (auto &a : PK) a.fill(-999999.0f);
    for (long long f = 0; f < nframes; f++) {
        const unsigned char *frame = (const unsigned char *)in + f * (long long)N * BPS;
        for (int rank = 0; rank < CS; rank++)
            for (int g = 0; g < C::UNITS; g++) {
                const int c0 = rank * C::COLS + g * CPC;
                for (int r = 0; r < N1; r++)  // the tensor-map box: N1 rows of CPC pairs
                    memcpy(tile.data() + (size_t)r * CPC * BPS, frame + ((size_t)r * 256 + c0) * BPS, (size_t)CPC * BPS);
                for (int tid = 0; tid < 256; tid++) {
                    const int col = tid % CPC, t = tid / CPC;
                    float wreg[16];
                    uint32_t raw[16];
                    FA::load_window(win.empty() ? nullptr : win.data(), c0 + col, t, wreg);
                    FA::load_raw_tile(tile.data(), col, t, raw);
                    FA::first(raw, wreg, U[tid].data());
                    C::a_scatter(xch.data(), col, t, U[tid].data());
                }
                for (int tid = 0; tid < 256; tid++) {
                    const int col = tid % CPC, t = tid / CPC, n2 = c0 + col;
                    cf twz[16], u[16];
                    C::load_col_tw(tz.data(), n2, t, twz);
                    C::a_second(xch.data(), tw1.data(), twz, col, t, u,
                                [&](int owner, int lrow, cf v) { Z[owner][(size_t)lrow * 256 + n2] = v; });
                }
            }
        float *out = rows + f * (long long)N;
        for (int rank = 0; rank < CS; rank++)
            for (int g = 0; g < C::UNITS; g++) {
                for (int tid = 0; tid < 256; tid++) {
                    const int row1 = tid / 16, t1 = tid % 16;
                    C::b_first(Z[rank].data() + (size_t)(16 * g + row1) * 256, t1, U[tid].data());
                }
                for (int tid = 0; tid < 256; tid++) C::b_scatter(xch.data(), tid / 16, tid % 16, U[tid].data());
                for (int tid = 0; tid < 256; tid++) {
                    const int row = tid % 16, tb = tid / 16, k1 = rank * C::ROWS + 16 * g + row;
                    float *pk = PK[((size_t)rank * C::UNITS + g) * 256 + tid].data();
                    if (f >= store_from)
                        C::template b_second<true, true>(xch.data(), tw256.data(), row, tb, k1, out, pk, bias);
                    else
                        C::template b_second<true, false>(xch.data(), tw256.data(), row, tb, k1, out, pk, bias);
                }
            }
    }
    if (peaks)
        for (int rank = 0; rank < CS; rank++)
            for (int g = 0; g < C::UNITS; g++)
                for (int tid = 0; tid < 256; tid++)
                    for (int c = 0; c < 16; c++)
                        peaks[C::bin_of(tid / 16, rank * C::ROWS + 16 * g + tid % 16, c)] = PK[((size_t)rank * C::UNITS + g) * 256 + tid][c];
    return 0;
}

template <int N1>
int emu_cluster_fmt(int in_fmt, const void *in, int window_kind, long long nframes, float *rows, float *peaks, long long store_from) {
    switch (in_fmt) {
        case FMT_S8: return emu_cluster<N1, FMT_S8>(in, window_kind, nframes, rows, peaks, store_from);
        case FMT_U8: return emu_cluster<N1, FMT_U8>(in, window_kind, nframes, rows, peaks, store_from);
        case FMT_S16LE: return emu_cluster<N1, FMT_S16LE>(in, window_kind, nframes, rows, peaks, store_from);
    }
    return -1;
}

extern "C" int emu_cluster_spectrum(int N, int in_fmt, int window_kind, const void *in, long long nframes, float *rows,
                                    float *peaks, long long store_from) {
    if (N == 65536) return emu_cluster_fmt<256>(in_fmt, in, window_kind, nframes, rows, peaks, store_from);
    if (N == 32768) return emu_cluster_fmt<128>(in_fmt, in, window_kind, nframes, rows, peaks, store_from);
    return -1;
}

// the dual-frame kernel's path (N = 256 .. 4096, integer formats, dB rows)
extern "C" int emu_spectrum2_avg(int N, int in_fmt, int window_kind, const void *in, long long nframes, float *rows,
                                 float *peaks, float *avg, int L, long long store_from) {
    return emu_dispatch(N, in_fmt, OUT_DB, window_kind, in, nullptr, nframes, rows, peaks, avg, L, true, store_from);
}

static int emu_dispatch(int N, int in_fmt, int out_kind, int window_kind, const void *in, const float *in_im,
                        long long nframes, float *rows, float *peaks, float *avg, int L, bool dual, long long store_from) {
    SpectrumParams p{};
    std::vector<float> win;
    if (window_kind >= 0) {
        win.resize(N);
        make_window(window_kind, N, win.data());
        p.win = win.data();
    }
    p.in = in;
    p.in_im = in_im;
    p.rows = rows;
    p.row0 = 0;
    p.row_step = 1;
    p.ring_rows = 0;
    p.row_stride = out_kind == OUT_CPLX ? 2LL * N : N;
    p.nframes = nframes;
    p.store_from = store_from;
    p.avg = avg;
    p.avg_len = L;
    p.avg_newest = nframes - 1;
    p.avg_dir = -1;
    p.avg_valid = nframes;
    if (dual) {
        switch (N) {
            case 256: return emu_two_fmt<256>(in_fmt, p, peaks);
            case 512: return emu_two_fmt<512>(in_fmt, p, peaks);
            case 1024: return emu_two_fmt<1024>(in_fmt, p, peaks);
            case 2048: return emu_two_fmt<2048>(in_fmt, p, peaks);
#ifndef RFA_R8  // (the radix-8 experiment changes Plan<4096>, which the dual-frame kernel does not follow)
            case 4096: return emu_two_fmt<4096>(in_fmt, p, peaks);
#endif
        }
        return -1;
    }
    switch (N) {
        case 16: return emu_size<16, 1>(in_fmt, out_kind, p, peaks);
        case 32: return emu_size<32, 1>(in_fmt, out_kind, p, peaks);
        case 64: return emu_size<64, 1>(in_fmt, out_kind, p, peaks);
        case 128: return emu_size<128, 1>(in_fmt, out_kind, p, peaks);
        case 256: return emu_size<256, 1>(in_fmt, out_kind, p, peaks);
        case 512: return emu_size<512, 1>(in_fmt, out_kind, p, peaks);
        case 1024: return emu_size<1024, 1>(in_fmt, out_kind, p, peaks);
        case 2048: return emu_size<2048, 1>(in_fmt, out_kind, p, peaks);
        case 4096: return emu_size<4096, 1>(in_fmt, out_kind, p, peaks);
        case 8192: return emu_size<8192, 1>(in_fmt, out_kind, p, peaks);
        case 16384: return emu_size<16384, 1>(in_fmt, out_kind, p, peaks);
        case 32768: return emu_size<16384, 2>(in_fmt, out_kind, p, peaks);
        case 65536: return emu_size<16384, 4>(in_fmt, out_kind, p, peaks);
    }
    return -1;
}

extern "C" int emu_spectrum(int N, int in_fmt, int out_kind, int window_kind, const void *in, const float *in_im,
                            long long nframes, float *rows, float *peaks) {
    return emu_spectrum_avg(N, in_fmt, out_kind, window_kind, in, in_im, nframes, rows, peaks, nullptr, 0);
}
