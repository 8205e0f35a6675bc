// spectrum_launch.h -- host-side interface of the fused spectrum kernels.
#pragma once
#include <cuda_runtime.h>

#include "spectrum_kernel.cuh"
#include "tuning.h"

namespace rfa {

struct SpectrumLaunch {
    int N;         // FFT size, power of two, 16 .. 65536
    int in_fmt;    // FMT_*
    int out_kind;  // OUT_DB / OUT_CPLX
    SpectrumParams p;
    cudaStream_t stream;
    int num_sms;
    int max_grid;  // 0 = size to the machine
    Tuning tune;   // the context's knobs (rfa_ctx_set_option)
};

// number of CTAs the launcher will use and the per-CTA frame slots (for peak_partial sizing)
cudaError_t spectrum_grid(int N, int in_fmt, int out_kind, long long nframes, int num_sms, int max_grid,
                          int *grid, int *slots_per_cta);
cudaError_t spectrum_launch(const SpectrumLaunch &L);

// instantiation groups (one translation unit each, see spectrum_inst.cu)
cudaError_t spectrum_group0(const SpectrumLaunch &L, bool query, int *grid, int *spc);
cudaError_t spectrum_group1(const SpectrumLaunch &L, bool query, int *grid, int *spc);
cudaError_t spectrum_group2(const SpectrumLaunch &L, bool query, int *grid, int *spc);
cudaError_t spectrum_group3(const SpectrumLaunch &L, bool query, int *grid, int *spc);

// four-step path for N = 32768 / 65536 (fourstep.cu): rows, peaks; the caller averages afterwards
struct FourStepLaunch {
    const cf *tw_n1;    // make_pass_twiddles(N / 256)
    const cf *tw_256;   // make_pass_twiddles(256)
    const cf *tw_n;     // make_twiddles(N)
    const cf *tz;       // [N / 256][256] column twiddles W_N^(n2 k1) of the cluster path
    cf *z;              // batch buffer
    long long z_bytes;
    unsigned int *sync;  // 2 * nframes counters for the fused launch (NULL: two kernels per batch)
};
bool fourstep_supported(int N, int in_fmt, int out_kind, const Tuning &tune);
int fourstep_launches(int N, long long nframes, long long z_bytes);  // of the most recent fourstep_launch on this thread
cudaError_t fourstep_launch(const SpectrumLaunch &L, const FourStepLaunch &fs);
// the same transform on thread-block clusters, the intermediate in distributed shared memory (fourstep_cluster.cuh): needs
// fs.tw_n1, fs.tw_256, fs.tz only.  cudaErrorNotSupported: this call cannot take that path (knob "cluster" = 0, unaligned
// input, no room for a cluster on the device) and the caller runs fourstep_launch instead.
cudaError_t fourstep_cluster_launch(const SpectrumLaunch &L, const FourStepLaunch &fs);

// small helper kernels (spectrum.cu)
void fill_f32(float *dst, size_t n, float v, cudaStream_t s);
// peaks[i] = max(accumulate ? peaks[i] : -999999, max_s partial[s][i])
void reduce_peaks(const float *partial, int slots, int N, float *peaks, bool accumulate, cudaStream_t s);
// avg[i] = (sum_{r=0..L} row(newest + r*dir)[i]) / (L+1), rows past `valid` count as -9999
void average_rows(const float *rows, long long newest, long long dir, long long ring_rows, long long row_stride,
                  long long valid, int L, int N, float *avg, cudaStream_t s);
// exponential average over frames first .. last (time order): a = a + alpha*(row - a), every operation rounded;
// from_state: start from avg[] instead of the first row
void ema_rows(const float *rows, long long row0, long long row_step, long long ring_rows, long long row_stride,
              long long first, long long last, float alpha, bool from_state, int N, float *avg, cudaStream_t s);
// mean dB over bins [b0, b1) of each of nrows rows (FftProcessor.kt:150-156)
void channel_strength(const float *rows, long long row0, long long row_step, long long ring_rows,
                      long long row_stride, long long nrows, int b0, int b1, float *out, cudaStream_t s);
// shift every ring row by `shift` bins, filling with -9999 (FftProcessor.kt:199-217)
void shift_rows(float *rows, long long nrows, long long row_stride, int N, int shift, cudaStream_t s);

}  // namespace rfa
