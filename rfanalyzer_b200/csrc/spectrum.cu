// spectrum.cu -- dispatch of the fused spectrum kernel plus the small reductions that
// follow it (peak merge, box-car time average, channel signal strength, history shift).
#include "spectrum_launch.h"

namespace rfa {

static cudaError_t dispatch(const SpectrumLaunch &L, bool query, int *grid, int *spc) {
    if (L.N <= 512) return spectrum_group0(L, query, grid, spc);
    if (L.N <= 2048) return spectrum_group1(L, query, grid, spc);
    if (L.N <= 8192) return spectrum_group2(L, query, grid, spc);
    return spectrum_group3(L, query, grid, spc);
}

cudaError_t spectrum_grid(int N, int in_fmt, int out_kind, long long nframes, int num_sms, int max_grid,
                          int *grid, int *slots_per_cta) {
    SpectrumLaunch L{};
    L.N = N;
    L.in_fmt = in_fmt;
    L.out_kind = out_kind;
    L.p.nframes = nframes;
    L.num_sms = num_sms;
    L.max_grid = max_grid;
    return dispatch(L, true, grid, slots_per_cta);
}

cudaError_t spectrum_launch(const SpectrumLaunch &L) { return dispatch(L, false, nullptr, nullptr); }

// ---------------------------------------------------------------------------
__global__ void fill_kernel(float *dst, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}
void fill_f32(float *dst, size_t n, float v, cudaStream_t s) {
    if (n == 0) return;
    size_t blocks = (n + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(dst, n, v);
}

// FftProcessor.kt:230-245: peaks[i] = max(peaks[i], row[i]); -999999f means "no peak".
__global__ void reduce_peaks_kernel(const float *partial, int slots, int N, float *peaks, bool accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float m = accumulate ? peaks[i] : -999999.0f;
    for (int s = 0; s < slots; s++) m = fmaxf(m, partial[(size_t)s * N + i]);
    peaks[i] = m;
}
void reduce_peaks(const float *partial, int slots, int N, float *peaks, bool accumulate, cudaStream_t s) {
    reduce_peaks_kernel<<<(N + 127) / 128, 128, 0, s>>>(partial, slots, N, peaks, accumulate);
}

// AnalyzerSurface.kt:683-684,710-714: the newest L+1 rows are summed newest -> oldest in
// float32 and divided by (L+1).  Rows that were never written hold -9999f (FftProcessor.kt:181).
__global__ void average_rows_kernel(const float *rows, long long newest, long long dir, long long ring_rows,
                                    long long row_stride, long long valid, int L, int N, float *avg) {
    // chained by programmatic dependent launch between the four-step row kernel and the next call's column kernel
    launch_dependents();
    grid_dependency_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float sum = 0.0f;
    for (int r = 0; r <= L; r++) {
        float v = -9999.0f;
        if (r < valid) {
            long long row = newest + (long long)r * dir;
            if (ring_rows > 0) {
                row %= ring_rows;
                if (row < 0) row += ring_rows;
            }
            v = rows[row * row_stride + i];
        }
        sum = __fadd_rn(sum, v);
    }
    avg[i] = __fdiv_rn(sum, (float)(L + 1));
}
void average_rows(const float *rows, long long newest, long long dir, long long ring_rows, long long row_stride,
                  long long valid, int L, int N, float *avg, cudaStream_t s) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((N + 127) / 128));
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, average_rows_kernel, rows, newest, dir, ring_rows, row_stride, valid, L, N, avg);
}

// RFA_AVG_EMA (an option the reference does not have; include/rfa_b200.h): one thread per bin walks the frames in time
// order, eight row loads in flight ahead of the dependent chain.  Joins the programmatic-dependent-launch chain of the
// spectrum kernels like average_rows_kernel.
__global__ void __launch_bounds__(128) ema_rows_kernel(const float *rows, long long row0, long long row_step, long long ring_rows,
                                                       long long row_stride, long long first, long long last, float alpha,
                                                       int from_state, int N, float *avg) {
    launch_dependents();
    grid_dependency_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    auto row_of = [&](long long f) {
        long long row = row0 + f * row_step;
        if (ring_rows > 0) {
            row %= ring_rows;
            if (row < 0) row += ring_rows;
        }
        return rows + row * row_stride + i;
    };
    long long f = first;
    float a;
    if (from_state) {
        a = avg[i];
    } else {
        a = __ldcg(row_of(f));
        f++;
    }
    for (; f + 8 <= last + 1; f += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldcg(row_of(f + k));
#pragma unroll
        for (int k = 0; k < 8; k++) a = __fadd_rn(a, __fmul_rn(alpha, __fsub_rn(v[k], a)));
    }
    for (; f <= last; f++) a = __fadd_rn(a, __fmul_rn(alpha, __fsub_rn(__ldcg(row_of(f)), a)));
    avg[i] = a;
}
void ema_rows(const float *rows, long long row0, long long row_step, long long ring_rows, long long row_stride,
              long long first, long long last, float alpha, bool from_state, int N, float *avg, cudaStream_t s) {
    if (last < first) return;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((N + 127) / 128));
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, ema_rows_kernel, rows, row0, row_step, ring_rows, row_stride, first, last, alpha,
                       from_state ? 1 : 0, N, avg);
}

// FftProcessor.kt:150-156: sequential float32 sum of mag[b0..b1) divided by the bin count.
// One warp per row keeps the reference's left-to-right order inside each lane's stripe
// only, so the result is within rounding of the reference, not bit-identical.
__global__ void channel_strength_kernel(const float *rows, long long row0, long long row_step, long long ring_rows,
                                        long long row_stride, long long nrows, int b0, int b1, float *out) {
    long long r = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    long long row = row0 + r * row_step;
    if (ring_rows > 0) {
        row %= ring_rows;
        if (row < 0) row += ring_rows;
    }
    const float *p = rows + row * row_stride;
    float sum = 0.0f;
    for (int i = b0 + lane; i < b1; i += 32) sum += p[i];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) out[r] = sum / (float)(b1 - b0);
}
void channel_strength(const float *rows, long long row0, long long row_step, long long ring_rows,
                      long long row_stride, long long nrows, int b0, int b1, float *out, cudaStream_t s) {
    if (nrows <= 0 || b1 <= b0) return;
    unsigned blocks = (unsigned)((nrows + 7) / 8);
    channel_strength_kernel<<<blocks, 256, 0, s>>>(rows, row0, row_step, ring_rows, row_stride, nrows, b0, b1, out);
}

// FftProcessor.kt:199-217: after a retune every history row moves by `shift` bins
// (shift < 0: left) and the vacated bins become -9999f.  One CTA per row moves it IN PLACE in
// chunks of 1024 bins, walking against the shift direction: a chunk is loaded completely
// before it is stored, and its destination only overlaps bins that were already moved.
__global__ void __launch_bounds__(256) shift_rows_kernel(float *rows, long long row_stride, int N, int shift) {
    float *p = rows + (long long)blockIdx.x * row_stride;
    const int chunks = (N + 1023) / 1024;
    for (int cc = 0; cc < chunks; cc++) {
        const int c = shift > 0 ? chunks - 1 - cc : cc;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int src = c * 1024 + j * 256 + threadIdx.x;
            v[j] = src < N ? p[src] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int src = c * 1024 + j * 256 + threadIdx.x, dst = src + shift;
            if (src < N && dst >= 0 && dst < N) p[dst] = v[j];
        }
        __syncthreads();
    }
    // vacated bins
    const int lo = shift > 0 ? 0 : (N + shift > 0 ? N + shift : 0);
    const int hi = shift > 0 ? (shift < N ? shift : N) : N;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) p[i] = -9999.0f;
}
void shift_rows(float *rows, long long nrows, long long row_stride, int N, int shift, cudaStream_t s) {
    if (nrows <= 0 || shift == 0) return;
    shift_rows_kernel<<<(unsigned)nrows, 256, 0, s>>>(rows, row_stride, N, shift);
}

}  // namespace rfa
