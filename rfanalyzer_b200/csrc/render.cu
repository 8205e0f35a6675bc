// render.cu -- waterfall / FFT-trace preprocessing on the device.
//
// Replaces the arithmetic of AnalyzerSurface.drawPreprocessing (A/ui/AnalyzerSurface.kt:646-734):
//   per pixel i of a waterfall row: mean of the row's bins j+start, j in [int(i*spp), (i+1)*spp)   (:703-713)
//   colour-map index int((avg - minDB) * scale), clamped to the map, -> ARGB through the map       (:726-727)
//   pixels outside (firstPixel, lastPixel-1) are black                                             (:728-731)
//   FFT trace: sum over the newest L+1 rows of those means, / (L+1)                                (:716-719)
//   peak trace: y = fftHeight - (mean of peaks over the pixel's bins - minDB) * dbWidth            (:714)
// Every sum runs in the reference's order in float32, so results are bit-identical to the JVM's.
// One thread per (row, pixel); rows and peaks are the device-resident ring the spectrum kernel fills,
// so a display client copies `width` pixels per row to the host instead of N floats.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace rfa {
namespace {

struct RenderArgs {
    const float *rows;       // ring [ring_rows][row_stride]
    long long row_stride;
    int ring_rows, n;        // rows in the ring, bins per row
    int newest;              // ring index of the newest row (currentRowIdx)
    int first_row, nrows;    // rowNumber range to render: [first_row, first_row + nrows), 0 = newest
    const float *peaks;      // [n] or nullptr
    int width;
    int start;               // first bin of the viewport (may be negative)
    float samples_per_px;
    int first_pixel, last_pixel;
    float min_db, scale, db_width, fft_height;
    const uint32_t *colormap;
    int colormap_size;
    uint32_t black;
    int avg_len;             // L
    uint32_t *argb;          // [ring_rows][width], row = ring index (like colorBuffer), or nullptr
    int *color_index;        // same layout, -1 outside the frame, or nullptr (parity tests)
    float *row_means;        // scratch [L+1][width]: horizontal means of the newest rows, or nullptr
    float *peaks_y;          // [width] or nullptr
};

__global__ void __launch_bounds__(256) render_rows_kernel(const RenderArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int row_number = a.first_row + blockIdx.y;
    if (i >= a.width) return;
    const int buffer_index = (a.newest + row_number) % a.ring_rows;
    const size_t px = (size_t)buffer_index * a.width + i;
    if (i >= a.first_pixel + 1 && i < a.last_pixel - 1) {
        const float *row = a.rows + (size_t)buffer_index * a.row_stride;
        const bool want_peaks = row_number == 0 && a.peaks != nullptr && a.peaks_y != nullptr;
        float avg = 0.0f, peak_avg = 0.0f;
        int counter = 0;
        int j = __float2int_rz(__fmul_rn((float)i, a.samples_per_px));  // (i * samplesPerPx).toInt()
        const float stop = __fmul_rn((float)(i + 1), a.samples_per_px);
        while ((float)j < stop && (j + a.start) < a.n) {
            avg = __fadd_rn(avg, row[j + a.start]);
            if (want_peaks) peak_avg = __fadd_rn(peak_avg, a.peaks[j + a.start]);
            counter++;
            j++;
        }
        avg = __fdiv_rn(avg, (float)counter);
        if (want_peaks)
            a.peaks_y[i] = __fsub_rn(a.fft_height, __fmul_rn(__fsub_rn(__fdiv_rn(peak_avg, (float)counter), a.min_db), a.db_width));
        if (a.row_means && row_number <= a.avg_len) a.row_means[(size_t)row_number * a.width + i] = avg;
        int idx = __float2int_rz(__fmul_rn(__fsub_rn(avg, a.min_db), a.scale));  // NaN -> 0, saturating: Java's (int)
        idx = idx < 0 ? 0 : (idx >= a.colormap_size ? a.colormap_size - 1 : idx);
        if (a.color_index) a.color_index[px] = idx;
        if (a.argb) a.argb[px] = a.colormap ? a.colormap[idx] : (uint32_t)idx;
    } else {
        if (a.color_index) a.color_index[px] = -1;
        if (a.argb) a.argb[px] = a.black;
        if (row_number == 0 && a.peaks != nullptr && a.peaks_y != nullptr) a.peaks_y[i] = -1.0f;
    }
}

// timeAverageSamples[i] = sum over rowNumber 0..L of the row means, in that order; / (L+1)
__global__ void __launch_bounds__(256) render_trace_kernel(const float *row_means, int width, int first_pixel,
                                                           int last_pixel, int avg_len, float *time_average) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    if (i >= first_pixel + 1 && i < last_pixel - 1) {
        float s = 0.0f;
        for (int r = 0; r <= avg_len; r++) s = __fadd_rn(s, row_means[(size_t)r * width + i]);
        time_average[i] = __fdiv_rn(s, (float)(avg_len + 1));
    } else {
        time_average[i] = __int_as_float(0x7fc00000);  // not drawn
    }
}

}  // namespace

cudaError_t render_launch(const RenderDesc &d, cudaStream_t st) {
    if (d.width <= 0 || d.nrows <= 0) return cudaSuccess;
    RenderArgs a{};
    a.rows = d.rows;
    a.row_stride = d.row_stride;
    a.ring_rows = d.ring_rows;
    a.n = d.n;
    a.newest = d.newest;
    a.first_row = d.first_row;
    a.nrows = d.nrows;
    a.peaks = d.peaks;
    a.width = d.width;
    a.start = d.start;
    a.samples_per_px = d.samples_per_px;
    a.first_pixel = d.first_pixel;
    a.last_pixel = d.last_pixel;
    a.min_db = d.min_db;
    a.scale = d.scale;
    a.db_width = d.db_width;
    a.fft_height = d.fft_height;
    a.colormap = d.colormap;
    a.colormap_size = d.colormap_size;
    a.black = d.black;
    a.avg_len = d.avg_len;
    a.argb = d.argb;
    a.color_index = d.color_index;
    a.row_means = d.row_means;
    a.peaks_y = d.peaks_y;
    dim3 grid((unsigned)((d.width + 255) / 256), (unsigned)d.nrows);
    render_rows_kernel<<<grid, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (d.time_average && d.row_means) {
        render_trace_kernel<<<(unsigned)((d.width + 255) / 256), 256, 0, st>>>(d.row_means, d.width, d.first_pixel,
                                                                                d.last_pixel, d.avg_len, d.time_average);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace rfa
