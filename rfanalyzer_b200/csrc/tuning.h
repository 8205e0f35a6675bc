// tuning.h -- the knobs of a context (rfa_ctx_set_option).  They replace the environment switches of round 1:
// nothing on a launch path reads the environment any more, and one process can hold contexts with different settings.
// The defaults ARE the product; the knobs exist for A/B timing runs and for the tests that compare two code paths.
// Options marked "lab" exist only in librfa_b200_lab.so (built with -DRFA_LAB, `make lab`), which also carries the
// experimental kernels that are slower than the defaults (spectrum2 / spectrum64 / pair / lean, the fused four-step
// launch, the residue split for integer input).
#pragma once

namespace rfa {

struct Tuning {
    int staged = 1;              // "staged": spectrum kernel reads raw IQ through TMA bulk copies (0: per-thread loads)
    int pdl = 1;                 // "pdl": programmatic dependent launch for back-to-back spectrum launches
    int max_grid = 0;            // "max_grid": cap of the persistent grid, 0 = size to the machine
    long long fs_batch_kib = 128 << 10;  // "fs_batch_kib": four-step intermediate per batch (default 128 MiB)
    int fs_tma = 1;              // "fs_tma": four-step column kernel loads raw IQ as 2-D tensor-map boxes
    int fs_ztma = 1;             // "fs_ztma": ... and stores the intermediate as one tensor-map store per frame
    int fs_pdl = 1;              // "fs_pdl": four-step kernels chained by programmatic dependent launch
    int chunk_kib = 8192;        // "chunk_kib": IQ bytes per chunk of the host-buffer H2D -> kernel -> D2H pipeline
    int rs_span = 0;             // "rs_span": samples staged per CTA by the tiled resampler, 0 = default
    // "cluster": N >= 32768 on thread-block clusters, the four-step intermediate in distributed shared memory
    // (fourstep_cluster.cuh): ONE launch and a third of the two-kernel path's HBM traffic, but measured 119 / 93 us
    // against 79 / 78 us per 2^24 samples (N = 65536 / 32768, profiles/r02d_cluster_experiments.txt) -- the remote
    // stores and the per-frame table loads queue behind the same load/store unit -- so the two-kernel path is the default
    int cluster = 0;
    // ---- lab builds only ----
    int kernel = 0;              // "kernel": 0 default, 1 dual-frame, 2 64x64, 3 anti-phase pair, 4 lean (N = 4096 variants)
    int fourstep = 1;            // "fourstep": 0 = residue-split kernel for N = 32768 / 65536 integer input
    int fs_fused = 0;            // "fs_fused": single cooperative producer/consumer launch
    long long fs_ring_kib = 32 << 10;  // "fs_ring_kib": its ring of intermediate frames
};

}  // namespace rfa
