// fourstep.cu -- launcher of the four-step spectrum path (fourstep_kernel.cuh) for N = 32768, 65536.
#include <stdlib.h>
#include <string.h>

#include "fourstep_cluster.cuh"
#include "fourstep_kernel.cuh"
#include "spectrum_launch.h"

namespace rfa {
namespace {

// K is a function-pointer TYPE (the same for every kernel of this file), so nothing here may be static
template <class K>
cudaError_t resident_ctas(K kern, size_t smem, int *occ) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 256, smem);
    if (e == cudaSuccess && *occ < 1) *occ = 1;
    return e;
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// the IQ bytes of a call as a [frames * N1][256 * bps] byte matrix, box = one column group of one frame
template <int N1>
bool make_input_map(const void *iq, long long nframes, int bps, CUtensorMap *tm, const Tuning &tune) {
    using G = GeomFS<N1>;
    if (!tune.fs_tma) return false;  // knob "fs_tma" = 0: per-thread loads (A/B timing runs, the fallback's test)
    EncodeTiledFn enc = encode_tiled();
    if (!enc || ((size_t)iq & 15) != 0 || nframes * N1 > 0x7FFFFFFFLL) return false;  // box coordinates are 32-bit signed
    const cuuint64_t dims[2] = {(cuuint64_t)256 * bps, (cuuint64_t)nframes * N1};
    const cuuint64_t strides[1] = {(cuuint64_t)256 * bps};
    const cuuint32_t box[2] = {(cuuint32_t)(G::CPC * bps), (cuuint32_t)N1};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(iq), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

thread_local int g_last_launches = 0;

#ifdef RFA_LAB
// One cooperative launch for the whole call (fourstep_fused_kernel).  Returns cudaErrorNotSupported when the
// configuration does not allow it; the caller then runs the two-kernel path.
template <int N1, int IN>
cudaError_t run_fused(const SpectrumLaunch &L, const FourStepLaunch &fs, const CUtensorMap &tmap) {
    using G = GeomFS<N1>;
    constexpr int BPS = in_elem_bytes<IN>();
    // Lab builds only (knob "fs_fused").  Measured on B200 it is correct but SLOWER than two kernels per batch
    // (2^24 samples: 119 / 125 us against 77 / 81 us, gpurun_out/fs_timing5.log), so it is not the default.
    if (!L.tune.fs_fused || !fs.sync || L.p.nframes > 0x7FFFFFFFLL) return cudaErrorNotSupported;
    auto kf = fourstep_fused_kernel<N1, IN>;
    const size_t smem = G::smem_fused(BPS);
    static thread_local int dev_done = -1, occ = 0, coop = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev_done != dev) {
        e = resident_ctas(kf, smem, &occ);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (e != cudaSuccess) return e;
        dev_done = dev;
    }
    // one producer and one consumer CTA per SM, rounded down to whole lanes
    const int lanes_a = L.num_sms / G::GROUPS_A, lanes_b = L.num_sms / G::GROUPS_B;
    if (!coop || occ < 2 || lanes_a < 1 || lanes_b < 1) return cudaErrorNotSupported;
    const long long ring_bytes = L.tune.fs_ring_kib << 10;  // Z ring size; default 32 MiB: a quarter of the L2
    long long ring = ring_bytes / ((long long)G::N * (long long)sizeof(cf));
    const long long min_ring = 2LL * lanes_b + lanes_a + 8;  // the consumers' look-ahead must stay inside the ring
    if (ring < min_ring) ring = min_ring;
    if (ring > L.p.nframes) ring = L.p.nframes;
    if (ring * (long long)G::N * (long long)sizeof(cf) > fs.z_bytes) return cudaErrorNotSupported;
    FourStepParams a{};
    a.p = L.p;
    a.tw_n1 = fs.tw_n1;
    a.tw_256 = fs.tw_256;
    a.tw_n = fs.tw_n;
    a.z = fs.z;
    a.frame0 = 0;
    a.nbatch = (int)L.p.nframes;
    a.ring = (int)ring;
    a.n_prod = lanes_a * G::GROUPS_A;
    a.col_done = fs.sync;
    a.row_done = fs.sync + L.p.nframes;
    e = cudaMemsetAsync(fs.sync, 0, 2 * (size_t)L.p.nframes * sizeof(unsigned int), L.stream);
    if (e != cudaSuccess) return e;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(256);
    cfg.gridDim = dim3((unsigned)(a.n_prod + lanes_b * G::GROUPS_B));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kf, a, tmap);
    if (e == cudaSuccess) g_last_launches = 1;
    return e;
}
#endif  // RFA_LAB

// the batch buffer Z as a [frames * N1][512] float matrix, box = the N1 x CPC points one column group finishes per frame
template <int N1>
bool make_z_map(cf *z, long long batch_frames, CUtensorMap *tm, const Tuning &tune) {
    using G = GeomFS<N1>;
    if (!tune.fs_ztma) return false;  // knob "fs_ztma" = 0: per-thread stores of Z (A/B timing runs)
    EncodeTiledFn enc = encode_tiled();
    if (!enc || ((size_t)z & 15) != 0 || batch_frames * N1 > 0x7FFFFFFFLL) return false;
    const cuuint64_t dims[2] = {512, (cuuint64_t)batch_frames * N1};
    const cuuint64_t strides[1] = {512 * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)(G::CPC * 2), (cuuint32_t)N1};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, z, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int N1, int IN>
cudaError_t run(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    using G = GeomFS<N1>;
    constexpr int BPS = in_elem_bytes<IN>();
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const bool staged = make_input_map<N1>(L.p.in, L.p.nframes, BPS, &tmap, L.tune);
    auto ka = staged ? fourstep_cols_kernel<N1, IN, true> : fourstep_cols_kernel<N1, IN, false>;
    auto kb = fourstep_rows_kernel<N1>;
    const size_t smem_a = staged ? G::smem_a_staged(BPS) : G::SMEM_A;
#ifdef RFA_LAB
    if (staged) {
        const cudaError_t ef = run_fused<N1, IN>(L, fs, tmap);
        if (ef != cudaErrorNotSupported) return ef;
    }
#endif
    g_last_launches = 0;
    static thread_local int dev_done[2] = {-1, -1}, occ_as[2] = {1, 1}, occ_b = 1;  // per instantiation <N1, IN>, per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev_done[staged] != dev) {
        e = resident_ctas(ka, smem_a, &occ_as[staged]);
        if (e != cudaSuccess) return e;
        e = resident_ctas(kb, G::SMEM_B, &occ_b);
        if (e != cudaSuccess) return e;
        dev_done[staged] = dev;
    }
    const int occ_a = occ_as[staged];
    const long long batch = fs.z_bytes / ((long long)G::N * (long long)sizeof(cf));
    if (batch < 1) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap tmap_z;
    memset(&tmap_z, 0, sizeof(tmap_z));
    // measured (gpurun_out/fs_timing6.log): 2.5 % faster for 8-bit IQ, 8 % slower for int16 IQ, whose raw tiles are
    // twice as large (the store tile then costs the second resident CTA its shared memory head-room)
    const bool ztma = staged && BPS == 2 && make_z_map<N1>(fs.z, batch, &tmap_z, L.tune);
    auto kz = fourstep_cols_ztma_kernel<N1, IN>;
    const size_t smem_z = G::smem_a_zstore(BPS);
    static thread_local int dev_done_z = -1, occ_z = 1;
    if (ztma && dev_done_z != dev) {
        e = resident_ctas(kz, smem_z, &occ_z);
        if (e != cudaSuccess) return e;
        dev_done_z = dev;
    }
    const bool pdl = L.tune.fs_pdl != 0;  // knob "fs_pdl" = 0: plain stream order (A/B timing runs)
    FourStepParams a{};
    a.p = L.p;
    a.tw_n1 = fs.tw_n1;
    a.tw_256 = fs.tw_256;
    a.tw_n = fs.tw_n;
    a.z = fs.z;
    a.ring = 0x7FFFFFFF;
    for (long long f0 = 0; f0 < L.p.nframes; f0 += batch) {
        const long long nb = L.p.nframes - f0 < batch ? L.p.nframes - f0 : batch;
        a.frame0 = f0;
        a.nbatch = (int)nb;
        long long lanes_a = (long long)L.num_sms * (ztma ? occ_z : occ_a) / G::GROUPS_A, lanes_b = (long long)L.num_sms * occ_b / G::GROUPS_B;
        if (lanes_a < 1) lanes_a = 1;
        if (lanes_b < 1) lanes_b = 1;
        if (lanes_a > nb) lanes_a = nb;
        if (lanes_b > nb) lanes_b = nb;
        // both launches carry the programmatic-serialization attribute: their prologues overlap the drain of
        // the kernel before them, griddepcontrol.wait in the kernels orders every access to shared data
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(256);
        cfg.stream = L.stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)(lanes_a * G::GROUPS_A));
        cfg.dynamicSmemBytes = ztma ? smem_z : smem_a;
        e = ztma ? cudaLaunchKernelEx(&cfg, kz, a, tmap, tmap_z) : cudaLaunchKernelEx(&cfg, ka, a, tmap);
        if (e != cudaSuccess) return e;
        cfg.gridDim = dim3((unsigned)(lanes_b * G::GROUPS_B));
        cfg.dynamicSmemBytes = G::SMEM_B;
        e = cudaLaunchKernelEx(&cfg, kb, a);
        if (e != cudaSuccess) return e;
        g_last_launches += 2;
    }
    return cudaSuccess;
}

// ---- cluster path (fourstep_cluster.cuh): one launch per call, Z in distributed shared memory ----
template <int N1, int IN>
cudaError_t run_cluster(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    using C = ClusterFS<N1, IN>;
    constexpr int BPS = in_elem_bytes<IN>();
    if (!L.tune.cluster || !fs.tz || L.p.nframes > 0x7FFFFFFFLL) return cudaErrorNotSupported;
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    Tuning tma_on = L.tune;
    tma_on.fs_tma = 1;  // the cluster kernel has no per-thread load path
    if (!make_input_map<N1>(L.p.in, L.p.nframes, BPS, &tmap, tma_on)) return cudaErrorNotSupported;
    auto kc = fourstep_cluster_kernel<N1, IN>;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = L.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static thread_local int dev_done = -1, max_clusters = 0;  // per instantiation <N1, IN>
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev_done != dev) {
        max_clusters = 0;
        if (cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) == cudaSuccess) {
            cfg.gridDim = dim3((unsigned)(C::CS * L.num_sms));
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kc, &cfg) != cudaSuccess) max_clusters = 0;
        }
        (void)cudaGetLastError();  // a device without room for this kernel takes the two-kernel path
        dev_done = dev;
    }
    if (max_clusters < 1) return cudaErrorNotSupported;
    long long clusters = max_clusters;
    if (L.max_grid > 0 && clusters > L.max_grid / C::CS) clusters = L.max_grid / C::CS > 0 ? L.max_grid / C::CS : 1;
    if (clusters > L.p.nframes) clusters = L.p.nframes;
    FourStepParams a{};
    a.p = L.p;
    a.tw_n1 = fs.tw_n1;
    a.tw_256 = fs.tw_256;
    a.tz = fs.tz;
    a.frame0 = 0;
    a.nbatch = (int)L.p.nframes;
    cfg.gridDim = dim3((unsigned)(clusters * C::CS));
    e = cudaLaunchKernelEx(&cfg, kc, a, tmap);
    if (e == cudaSuccess) g_last_launches = 1;
    return e;
}

template <int N1>
cudaError_t run_cluster_fmt(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    switch (L.in_fmt) {
        case FMT_S8: return run_cluster<N1, FMT_S8>(L, fs);
        case FMT_U8: return run_cluster<N1, FMT_U8>(L, fs);
        case FMT_S16LE: return run_cluster<N1, FMT_S16LE>(L, fs);
    }
    return cudaErrorInvalidValue;
}

template <int N1>
cudaError_t run_fmt(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    switch (L.in_fmt) {
        case FMT_S8: return run<N1, FMT_S8>(L, fs);
        case FMT_U8: return run<N1, FMT_U8>(L, fs);
        case FMT_S16LE: return run<N1, FMT_S16LE>(L, fs);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

bool fourstep_supported(int N, int in_fmt, int out_kind, const Tuning &tune) {
#ifdef RFA_LAB
    if (!tune.fourstep) return false;  // lab knob "fourstep" = 0: the residue-split kernel, a second factorisation
#else
    (void)tune;
#endif
    return (N == 32768 || N == 65536) && out_kind == OUT_DB && (in_fmt == FMT_S8 || in_fmt == FMT_U8 || in_fmt == FMT_S16LE);
}

int fourstep_launches(int, long long, long long) { return g_last_launches; }

cudaError_t fourstep_cluster_launch(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    if (L.p.nframes <= 0) return cudaSuccess;
    return L.N == 65536 ? run_cluster_fmt<256>(L, fs) : run_cluster_fmt<128>(L, fs);
}

cudaError_t fourstep_launch(const SpectrumLaunch &L, const FourStepLaunch &fs) {
    if (L.p.nframes <= 0) return cudaSuccess;
    return L.N == 65536 ? run_fmt<256>(L, fs) : run_fmt<128>(L, fs);
}

}  // namespace rfa
