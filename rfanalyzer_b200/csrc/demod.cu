// demod.cu -- K6 demodulators: Demodulator.demodulateFM / AM / SSB / CW and the
// volume scaling of Demodulator.run (A/analyzer/Demodulator.kt:184-187, :251-403).
//
// FM is a stream operator with a one-sample memory.  AM, SSB and CW normalise PER PACKET with
// a gain that follows lastMax <- max(0.95f*lastMax, max(packet)): packetisation is part of
// the reference's semantics, so these kernels take the packet boundaries (segment offsets in
// the demodulator's input/output domain) and run (i) a per-packet reduction, (ii) a one-thread
// scan over packets for the gain recurrence, (iii) an element-wise normalisation.
// SUM_EXACT reproduces the JVM's float32 arithmetic bit for bit (separately rounded
// products, sequential per-packet mean, double-precision atan2); SUM_FMA is the fast variant.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"
#include "pdl.h"

namespace rfa {
namespace {

// Demodulator.kt:262-270
template <bool EXACT>
__global__ void __launch_bounds__(256) fm_kernel(const float *__restrict__ re, const float *__restrict__ im,
                                                 long long n, const float *__restrict__ carry, float gain,
                                                 float volume, float *__restrict__ out) {
    pdl_enter();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float r = re[i], q = im[i];
        const float pr = i ? re[i - 1] : carry[0], pq = i ? im[i - 1] : carry[1];
        float a, b;
        if (EXACT) {
            a = __fadd_rn(__fmul_rn(r, pr), __fmul_rn(q, pq));
            b = __fsub_rn(__fmul_rn(q, pr), __fmul_rn(r, pq));
            const float ph = (float)atan2((double)b, (double)a);
            out[i] = __fmul_rn(__fmul_rn(gain, ph), volume);
        } else {
            a = fmaf(r, pr, q * pq);
            b = fmaf(q, pr, -(r * pq));
            out[i] = gain * atan2f(b, a) * volume;
        }
    }
}

__global__ void carry_kernel(const float *re, const float *im, long long n, float *carry) {
    pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0 && n > 0) {
        carry[0] = re[n - 1];
        carry[1] = im[n - 1];
    }
}

// Demodulator.kt:293-297: power = re*re + im*im (each product rounded)
__global__ void __launch_bounds__(256) power_kernel(const float *__restrict__ re, const float *__restrict__ im,
                                                    long long n, float *__restrict__ out) {
    pdl_enter();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fadd_rn(__fmul_rn(re[i], re[i]), __fmul_rn(im[i], im[i]));
}

// per-packet sum and max of x over [off[p], off[p+1]).  EXACT: one thread adds the packet in
// index order (the reference's float32 running sum); otherwise a block-wide tree.
template <bool EXACT>
__global__ void __launch_bounds__(256) packet_stats_kernel(const float *__restrict__ x, const long long *__restrict__ off,
                                                           int npackets, float *__restrict__ sum,
                                                           float *__restrict__ mx) {
    pdl_enter();
    if (EXACT) {
        const int p = blockIdx.x * blockDim.x + threadIdx.x;
        if (p >= npackets) return;
        float s = 0.0f, m = -INFINITY;
        for (long long i = off[p]; i < off[p + 1]; i++) {
            s = __fadd_rn(s, x[i]);
            m = fmaxf(m, x[i]);
        }
        sum[p] = s;
        mx[p] = m;
    } else {
        const int p = blockIdx.x;
        __shared__ float ss[256], sm[256];
        float s = 0.0f, m = -INFINITY;
        for (long long i = off[p] + threadIdx.x; i < off[p + 1]; i += blockDim.x) {
            s += x[i];
            m = fmaxf(m, x[i]);
        }
        ss[threadIdx.x] = s;
        sm[threadIdx.x] = m;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) {
                ss[threadIdx.x] += ss[threadIdx.x + o];
                sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            sum[p] = ss[0];
            mx[p] = sm[0];
        }
    }
}

// Demodulator.kt:290,296,299-302 (AM) and :347-355 (SSB/CW): the AGC recurrence over packets.
// state[0] = lastMax carried between calls.  mean[p] only matters for AM.
// One CTA: the packet maxima go through shared memory in blocks of 1024, thread 0 runs the
// recurrence on them (the only sequential part, a few cycles per packet), everybody divides.
__global__ void __launch_bounds__(256) agc_scan_kernel(const long long *off, int npackets, const float *sum,
                                                       const float *mx, float *state, float *gain, float *mean) {
    pdl_enter();
    __shared__ float s_last[1024];
    __shared__ float s_carry;
    if (threadIdx.x == 0) s_carry = state[0];
    for (int p0 = 0; p0 < npackets; p0 += 1024) {
        const int np = npackets - p0 < 1024 ? npackets - p0 : 1024;
        for (int i = threadIdx.x; i < np; i += blockDim.x) s_last[i] = mx[p0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            float last = s_carry;
            for (int i = 0; i < np; i++) {
                last = __fmul_rn(last, (float)0.95);
                const float m = s_last[i];
                if (m > last) last = m;
                s_last[i] = last;
            }
            s_carry = last;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const int p = p0 + i;
            gain[p] = __fdiv_rn(0.75f, s_last[i]);
            const long long n = off[p + 1] - off[p];
            mean[p] = __fdiv_rn(sum[p], (float)n);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) state[0] = s_carry;
}

// out = (x - mean[p]) * gain[p] * volume (AM) or x * gain[p] * volume (SSB/CW); one block row per packet
template <bool SUBTRACT_MEAN>
__global__ void __launch_bounds__(256) normalise_kernel(float *__restrict__ x, const long long *__restrict__ off,
                                                        const float *__restrict__ gain, const float *__restrict__ mean,
                                                        float volume, int npackets) {
    pdl_enter();
    // gridDim.y is capped at 65535 by CUDA: a block row strides over the packets
    for (int p = blockIdx.y; p < npackets; p += gridDim.y) {
        const float g = gain[p], m = SUBTRACT_MEAN ? mean[p] : 0.0f;
        const long long b = off[p], e = off[p + 1];
        for (long long i = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
            float v = x[i];
            if (SUBTRACT_MEAN) v = __fsub_rn(v, m);
            x[i] = __fmul_rn(__fmul_rn(v, g), volume);
        }
    }
}

unsigned blocks_for(long long n, int num_sms) {
    long long b = (n + 255) / 256, cap = (long long)num_sms * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

cudaError_t demod_fm_launch(const float *re, const float *im, long long n, float *carry, float gain, float volume,
                            float *out, bool exact, int num_sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (exact)
        pdl_launch(fm_kernel<true>, blocks_for(n, num_sms), 256, 0, st, re, im, n, carry, gain, volume, out);
    else
        pdl_launch(fm_kernel<false>, blocks_for(n, num_sms), 256, 0, st, re, im, n, carry, gain, volume, out);
    pdl_launch(carry_kernel, 1, 32, 0, st, re, im, n, carry);  // Demodulator.kt:271-272
    return cudaGetLastError();
}

cudaError_t demod_power_launch(const float *re, const float *im, long long n, float *out, int num_sms,
                               cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    pdl_launch(power_kernel, blocks_for(n, num_sms), 256, 0, st, re, im, n, out);
    return cudaGetLastError();
}

cudaError_t agc_launch(float *x, const long long *off, int npackets, long long max_packet, bool subtract_mean,
                       float *state, float *scratch /* 4*npackets floats */, float volume, bool exact, int num_sms,
                       cudaStream_t st) {
    if (npackets <= 0) return cudaSuccess;
    float *sum = scratch, *mx = scratch + npackets, *gain = scratch + 2 * (size_t)npackets,
          *mean = scratch + 3 * (size_t)npackets;
    if (exact)
        pdl_launch(packet_stats_kernel<true>, (npackets + 63) / 64, 64, 0, st, x, off, npackets, sum, mx);
    else
        pdl_launch(packet_stats_kernel<false>, npackets, 256, 0, st, x, off, npackets, sum, mx);
    pdl_launch(agc_scan_kernel, 1, 256, 0, st, off, npackets, sum, mx, state, gain, mean);
    unsigned bx = (unsigned)((max_packet + 255) / 256);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid(bx, (unsigned)(npackets < 65535 ? npackets : 65535));
    if (subtract_mean)
        pdl_launch(normalise_kernel<true>, grid, 256, 0, st, x, off, gain, mean, volume, npackets);
    else
        pdl_launch(normalise_kernel<false>, grid, 256, 0, st, x, off, gain, mean, volume, npackets);
    return cudaGetLastError();
}

}  // namespace rfa
