// spectrum_inst.cu -- instantiations of the fused spectrum kernel, compiled once per
// size group (-DRFA_GROUP=0..3) so the groups build in parallel.
#include <stdlib.h>

#include "spectrum_launch.h"

namespace rfa {
namespace {

template <int NL, int S, int IN, int OUT>
cudaError_t launch_one(const SpectrumLaunch &L, bool query, int *grid_out, int *spc_out) {
    using G = Geom<NL>;
    constexpr size_t SMEM = SpectrumFrame<NL, S, IN, OUT>::SMEM_BYTES;
    auto kern = spectrum_kernel<NL, S, IN, OUT>;
    static bool configured = false;
    cudaError_t err;
    if (!configured) {
        if (SMEM > 48 * 1024) {
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            if (err != cudaSuccess) return err;
        }
        configured = true;
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::CTA, SMEM);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long need = ((L.p.nframes + G::FPC - 1) / G::FPC) * S;
    long long cap = (long long)L.num_sms * occ;
    static const int maxgrid_env = [] {  // tuning runs only
        const char *e = getenv("RFA_MAXGRID");
        return e ? atoi(e) : 0;
    }();
    if (maxgrid_env > 0 && maxgrid_env < cap) cap = maxgrid_env;
    if (L.max_grid > 0 && L.max_grid < cap) cap = L.max_grid;
    cap -= cap % S;
    if (cap < S) cap = S;
    long long grid = need < cap ? need : cap;
    if (grid < S) grid = S;
    if (grid_out) *grid_out = (int)grid;
    if (spc_out) *spc_out = G::FPC;
    if (query) return cudaSuccess;
    kern<<<(unsigned)grid, G::CTA, SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}

template <int NL, int S>
cudaError_t launch_size(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    if (L.out_kind == OUT_CPLX) {
        if (L.in_fmt == FMT_CF32) return launch_one<NL, S, FMT_CF32, OUT_CPLX>(L, query, grid, spc);
        return cudaErrorInvalidValue;
    }
    switch (L.in_fmt) {
        case FMT_S8: return launch_one<NL, S, FMT_S8, OUT_DB>(L, query, grid, spc);
        case FMT_U8: return launch_one<NL, S, FMT_U8, OUT_DB>(L, query, grid, spc);
        case FMT_S16LE: return launch_one<NL, S, FMT_S16LE, OUT_DB>(L, query, grid, spc);
        case FMT_CF32: return launch_one<NL, S, FMT_CF32, OUT_DB>(L, query, grid, spc);
        case FMT_PF32: return launch_one<NL, S, FMT_PF32, OUT_DB>(L, query, grid, spc);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

#if RFA_GROUP == 0
cudaError_t spectrum_group0(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    switch (L.N) {
        case 16: return launch_size<16, 1>(L, query, grid, spc);
        case 32: return launch_size<32, 1>(L, query, grid, spc);
        case 64: return launch_size<64, 1>(L, query, grid, spc);
        case 128: return launch_size<128, 1>(L, query, grid, spc);
        case 256: return launch_size<256, 1>(L, query, grid, spc);
        case 512: return launch_size<512, 1>(L, query, grid, spc);
    }
    return cudaErrorInvalidValue;
}
#elif RFA_GROUP == 1
cudaError_t spectrum_group1(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    switch (L.N) {
        case 1024: return launch_size<1024, 1>(L, query, grid, spc);
        case 2048: return launch_size<2048, 1>(L, query, grid, spc);
    }
    return cudaErrorInvalidValue;
}
#elif RFA_GROUP == 2
cudaError_t spectrum_group2(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    switch (L.N) {
        case 4096: return launch_size<4096, 1>(L, query, grid, spc);
        case 8192: return launch_size<8192, 1>(L, query, grid, spc);
    }
    return cudaErrorInvalidValue;
}
#elif RFA_GROUP == 3
cudaError_t spectrum_group3(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    switch (L.N) {
        case 16384: return launch_size<16384, 1>(L, query, grid, spc);
        case 32768: return launch_size<16384, 2>(L, query, grid, spc);
        case 65536: return launch_size<16384, 4>(L, query, grid, spc);
    }
    return cudaErrorInvalidValue;
}
#endif

}  // namespace rfa
