// spectrum_inst.cu -- instantiations of the fused spectrum kernel, compiled once per
// size group (-DRFA_GROUP=0..3) so the groups build in parallel.
//
// RFA_LAB (librfa_b200_lab.so, `make lab`) adds the experimental N = 4096 kernels that measured slower than the default
// (dual-frame, 64 x 64, anti-phase pair, lean) behind the context knob "kernel", and the residue split for integer input.
#include "device_once.h"
#include "spectrum_launch.h"
#ifdef RFA_LAB
#include "spectrum_pair_kernel.cuh"
#include "spectrum_lean_kernel.cuh"
#include "spectrum2_kernel.cuh"
#include "spectrum64_kernel.cuh"
#endif

namespace rfa {
namespace {

template <int NL, int S, int IN, int OUT, bool STAGED = false>
cudaError_t launch_one(const SpectrumLaunch &L, bool query, int *grid_out, int *spc_out) {
    using G = Geom<NL>;
    constexpr size_t SMEM = STAGED ? staged_offset<NL, S, IN, OUT>() + staged_bytes<NL, IN>()
                                   : SpectrumFrame<NL, S, IN, OUT>::SMEM_BYTES;
    auto kern = spectrum_kernel<NL, S, IN, OUT, STAGED>;
    static DeviceOnce once;  // per instantiation
    int dev = 0;
    cudaError_t err;
    if (once.pending(&dev)) {
        if (SMEM > 48 * 1024) {
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            if (err != cudaSuccess) return err;
        }
        once.done(dev);
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::CTA, SMEM);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long need = ((L.p.nframes + G::FPC - 1) / G::FPC) * S;
    if (need / S > 0x7FFF0000LL) return cudaErrorInvalidValue;  // chunk indices are 32-bit in the kernel
    long long cap = (long long)L.num_sms * occ;
    if (L.tune.max_grid > 0 && L.tune.max_grid < cap) cap = L.tune.max_grid;  // knob "max_grid" (occupancy experiments)
    if (L.max_grid > 0 && L.max_grid < cap) cap = L.max_grid;
    // the averaging CTA (spectrum_kernel.cuh: average_cta) takes one resident slot
    const bool avg_cta = OUT == OUT_DB && L.p.avg != nullptr;
    if (avg_cta && cap > S) cap -= 1;
    cap -= cap % S;
    if (cap < S) cap = S;
    long long grid = need < cap ? need : cap;
    if (grid < S) grid = S;
    if (grid_out) *grid_out = (int)grid;
    if (spc_out) *spc_out = G::FPC;
    if (query) return cudaSuccess;
    if (avg_cta) grid += 1;
    {
        // back-to-back launches overlap: programmatic stream serialization + griddepcontrol in the kernel (knob "pdl" = 0: off)
        if (L.tune.pdl) {
            SpectrumParams pp = L.p;
            pp.pdl = 1;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(G::CTA);
            cfg.dynamicSmemBytes = SMEM;
            cfg.stream = L.stream;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            return cudaLaunchKernelEx(&cfg, kern, pp);
        }
    }
    kern<<<(unsigned)grid, G::CTA, SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}

// TMA-staged input (spectrum_kernel.cuh, STAGED): N = 1024 .. 4096, integer formats, 16-byte aligned IQ.
// The knob "staged" = 0 switches it off (A/B timing runs).
static bool staged_enabled(const SpectrumLaunch &L) { return L.tune.staged && ((size_t)L.p.in & 15) == 0; }

#ifdef RFA_LAB
// dual-frame kernel (spectrum2_kernel.cuh): one 256-thread CTA per SM, two frames per thread
template <int NL, int IN>
cudaError_t launch_two(const SpectrumLaunch &L, bool query, int *grid_out, int *spc_out) {
    using G = Geom2<NL>;
    constexpr size_t SMEM = SpectrumFrame2<NL, IN>::SMEM_BYTES;
    auto kern = spectrum2_kernel<NL, IN>;
    static DeviceOnce once;
    int dev = 0;
    cudaError_t err;
    if (once.pending(&dev)) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (err != cudaSuccess) return err;
        once.done(dev);
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::CTA, SMEM);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    const long long npairs = (L.p.nframes + 1) / 2;
    long long need = (npairs + G::FPC - 1) / G::FPC;
    if (need > 0x7FFF0000LL) return cudaErrorInvalidValue;  // chunk indices are 32-bit in the kernel
    long long cap = (long long)L.num_sms * occ;
    if (L.max_grid > 0 && L.max_grid < cap) cap = L.max_grid;
    const bool avg_cta = L.p.avg != nullptr;  // the averaging CTA takes one resident slot
    if (avg_cta && cap > 1) cap -= 1;
    long long grid = need < cap ? need : cap;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    if (spc_out) *spc_out = 2 * G::FPC;
    if (query) return cudaSuccess;
    if (avg_cta) grid += 1;
    kern<<<(unsigned)grid, G::CTA, SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}

// two-pass 64 x 64 kernel for N = 4096 (spectrum64_kernel.cuh, experiment: RFA_K64=1, no time average)
template <int IN>
cudaError_t launch_64(const SpectrumLaunch &L) {
    using G = Geom64;
    auto kern = spectrum64_kernel<IN>;
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (err != cudaSuccess) return err;
        once.done(dev);
    }
    long long need = (L.p.nframes + G::SLOTS - 1) / G::SLOTS;
    long long grid = need < L.num_sms ? need : L.num_sms;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, G::CTA, G::SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}
// anti-phase pair kernel for N = 4096 (spectrum_pair_kernel.cuh): one 512-thread CTA per SM, two frames one segment apart
template <int IN>
cudaError_t launch_pair(const SpectrumLaunch &L, bool query, int *grid_out, int *spc_out) {
    using G = GeomPair;
    constexpr size_t SMEM = G::smem(in_elem_bytes<IN>());
    auto kern = spectrum_pair_kernel<IN>;
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (err != cudaSuccess) return err;
        once.done(dev);
    }
    const long long need = (L.p.nframes + G::SLOTS - 1) / G::SLOTS;
    if (need > 0x7FFF0000LL) return cudaErrorInvalidValue;  // chunk indices are 32-bit in the kernel
    const bool avg_cta = L.p.avg != nullptr;  // the averaging CTA takes one SM
    long long cap = L.num_sms;
    if (L.max_grid > 0 && L.max_grid < cap) cap = L.max_grid;
    if (avg_cta && cap > 1) cap -= 1;
    long long grid = need < cap ? need : cap;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    if (spc_out) *spc_out = G::SLOTS;
    if (query) return cudaSuccess;
    if (avg_cta) grid += 1;
    kern<<<(unsigned)grid, G::CTA, SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}
// three-CTAs-per-SM kernel for N = 4096, 8-bit IQ (spectrum_lean_kernel.cuh)
template <int IN>
cudaError_t launch_lean(const SpectrumLaunch &L, bool query, int *grid_out, int *spc_out) {
    using G = GeomLean;
    constexpr size_t SMEM = G::smem(in_elem_bytes<IN>());
    auto kern = spectrum_lean_kernel<IN>;
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (err != cudaSuccess) return err;
        once.done(dev);
    }
    int occ = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::CTA, SMEM);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    const long long need = L.p.nframes;
    if (need > 0x7FFF0000LL) return cudaErrorInvalidValue;  // chunk indices are 32-bit in the kernel
    const bool avg_cta = L.p.avg != nullptr;  // the averaging CTA takes one resident slot
    long long cap = (long long)L.num_sms * occ;
    if (L.max_grid > 0 && L.max_grid < cap) cap = L.max_grid;
    if (avg_cta && cap > 1) cap -= 1;
    long long grid = need < cap ? need : cap;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    if (spc_out) *spc_out = 1;
    if (query) return cudaSuccess;
    if (avg_cta) grid += 1;
    kern<<<(unsigned)grid, G::CTA, SMEM, L.stream>>>(L.p);
    return cudaGetLastError();
}
#endif  // RFA_LAB

template <int NL, int S>
cudaError_t launch_size(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
#ifdef RFA_LAB
    const bool al16 = ((size_t)L.p.in & 15) == 0;
    if constexpr (S == 1 && NL == 4096) {
        if (L.out_kind == OUT_DB && !query && L.tune.kernel == 2 && L.p.avg == nullptr && L.p.twN != nullptr) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_64<FMT_S8>(L);
                case FMT_U8: return launch_64<FMT_U8>(L);
                case FMT_S16LE: return launch_64<FMT_S16LE>(L);
            }
        }
        if (L.out_kind == OUT_DB && L.tune.kernel == 4 && al16) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_lean<FMT_S8>(L, query, grid, spc);
                case FMT_U8: return launch_lean<FMT_U8>(L, query, grid, spc);
            }
        }
        if (L.out_kind == OUT_DB && L.tune.kernel == 3 && al16) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_pair<FMT_S8>(L, query, grid, spc);
                case FMT_U8: return launch_pair<FMT_U8>(L, query, grid, spc);
                case FMT_S16LE: return launch_pair<FMT_S16LE>(L, query, grid, spc);
            }
        }
    }
    if constexpr (S == 1 && NL >= 256 && NL <= 4096) {
        if (L.out_kind == OUT_DB && L.tune.kernel == 1) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_two<NL, FMT_S8>(L, query, grid, spc);
                case FMT_U8: return launch_two<NL, FMT_U8>(L, query, grid, spc);
                case FMT_S16LE: return launch_two<NL, FMT_S16LE>(L, query, grid, spc);
            }
        }
    }
#endif  // RFA_LAB
    if (L.out_kind == OUT_CPLX) {
        if (L.in_fmt == FMT_CF32) return launch_one<NL, S, FMT_CF32, OUT_CPLX>(L, query, grid, spc);
        return cudaErrorInvalidValue;
    }
    if constexpr (S == 1 && NL >= 1024 && NL <= 8192) {
        if (staged_enabled(L)) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_one<NL, S, FMT_S8, OUT_DB, true>(L, query, grid, spc);
                case FMT_U8: return launch_one<NL, S, FMT_U8, OUT_DB, true>(L, query, grid, spc);
                // (at 8192 two 32 KB chunk buffers no longer fit beside the frames and tables: one buffer, stage_buffers();
                // with two 256-thread CTAs per SM the per-thread loads measured faster there: 55.2 against 59.2 us)
                case FMT_S16LE:
#ifndef RFA_8192_T512
                    if (NL == 8192) break;
#endif
                    return launch_one<NL, S, FMT_S16LE, OUT_DB, true>(L, query, grid, spc);
            }
        }
    }
    if constexpr (S == 1 && NL == 16384) {  // two 32 KB chunk buffers (8-bit IQ) or one of 64 KB (int16) beside the 139 KB exchange frame
        if (staged_enabled(L)) {
            switch (L.in_fmt) {
                case FMT_S8: return launch_one<NL, S, FMT_S8, OUT_DB, true>(L, query, grid, spc);
                case FMT_U8: return launch_one<NL, S, FMT_U8, OUT_DB, true>(L, query, grid, spc);
                case FMT_S16LE: return launch_one<NL, S, FMT_S16LE, OUT_DB, true>(L, query, grid, spc);
            }
        }
    }
#ifndef RFA_LAB
    // the residue split (S > 1) serves float input only (rfa_fft_*, N = 32768 / 65536); integer IQ of those sizes
    // goes through the cluster / four-step paths, its residue-split instantiations live in the lab build
    if constexpr (S > 1)
        if (L.in_fmt == FMT_S8 || L.in_fmt == FMT_U8 || L.in_fmt == FMT_S16LE) return cudaErrorNotSupported;
#endif
    switch (L.in_fmt) {
#if defined(RFA_LAB)
        case FMT_S8: return launch_one<NL, S, FMT_S8, OUT_DB>(L, query, grid, spc);
        case FMT_U8: return launch_one<NL, S, FMT_U8, OUT_DB>(L, query, grid, spc);
        case FMT_S16LE: return launch_one<NL, S, FMT_S16LE, OUT_DB>(L, query, grid, spc);
#else
        case FMT_S8: if constexpr (S == 1) return launch_one<NL, S, FMT_S8, OUT_DB>(L, query, grid, spc); break;
        case FMT_U8: if constexpr (S == 1) return launch_one<NL, S, FMT_U8, OUT_DB>(L, query, grid, spc); break;
        case FMT_S16LE: if constexpr (S == 1) return launch_one<NL, S, FMT_S16LE, OUT_DB>(L, query, grid, spc); break;
#endif
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

#if defined(RFA_TRACE) && RFA_GROUP == 2
// tuning builds: copy the phase stamps of the last N = 4096 / 8192 launch to the host
extern "C" int rfa_debug_trace(long long *out, int ctas, int slots) {
    static long long host[512][RFA_TRACE_SLOTS];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(host, rfa::g_trace, sizeof(host)) != cudaSuccess) return -2;
    for (int c = 0; c < ctas && c < 512; c++)
        for (int k = 0; k < slots && k < RFA_TRACE_SLOTS; k++) out[(size_t)c * slots + k] = host[c][k];
    return 0;
}
#endif
