// spectrum64_kernel.cuh -- two-pass variant of the fused IQ -> spectrum kernel for N = 4096 (EXPERIMENT,
// selected with RFA_K64=1).
//
// Same computation and reference lines as spectrum_kernel.cuh.  N = 4096 is split 64 x 64: a thread
// keeps 64 points in registers, a frame belongs to 64 threads (two warps), and there is ONE
// shared-memory exchange per frame instead of two, synchronised by a 64-thread named barrier instead
// of a CTA barrier.  Four frame slots per 256-thread CTA run independently (one CTA per SM, up to 255
// registers per thread); window taps and the 64 x 64 twiddle table are shared by the slots in shared
// memory, the 64 running peaks of a thread stay in registers.
// Measured against spectrum_kernel before it replaces anything: see DESIGN.md 4.1.
#pragma once
#include "spectrum_kernel.cuh"

namespace rfa {

// 64-point DFT in registers.  Storage: input point n lives in u[(n & 3) * 16 + (n >> 2)], natural-order
// output c in u[perm(c)].  64 = 4 x 16: DFT16 over n1 for each n2 (n = n2 + 4*n1), twiddle
// W64^(n2*k1), DFT4 over n2 for each k1 -> X[k1 + 16*k2].
template <>
struct Dft<64> {
    static RFA_CX int in_slot(int n) { return (n & 3) * 16 + (n >> 2); }
    static RFA_CX int perm(int c) { return 16 * (c >> 4) + Dft<16>::perm(c & 15); }
    template <class V>
    static RFA_HD void run(V *u) {
#pragma unroll
        for (int n2 = 0; n2 < 4; n2++) Dft<16>::run(u + 16 * n2);
#pragma unroll
        for (int n2 = 1; n2 < 4; n2++)
#pragma unroll
            for (int k1 = 1; k1 < 16; k1++) {
                const int e = n2 * k1;  // W64^e, 1 <= e <= 45
                V &v = u[16 * n2 + Dft<16>::perm(k1)];
                if (e == 16)
                    v = mul_mj(v);
                else if (e == 32)
                    v = csub(V{}, v);
                else
                    v = cmul(v, w64(e));
            }
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            const int p = Dft<16>::perm(k1);
            bfly4(u[p], u[16 + p], u[32 + p], u[48 + p]);
        }
    }
    // exp(-2*pi*i*e/64), correctly rounded constants
    static RFA_HD cf w64(int e) {
        const float c[17] = {1.0f,
                             0.99518472667219688624f,
                             0.98078528040323044913f,
                             0.95694033573220886494f,
                             0.92387953251128675613f,
                             0.88192126434835502971f,
                             0.83146961230254523708f,
                             0.77301045336273696081f,
                             0.70710678118654752440f,
                             0.63439328416364549822f,
                             0.55557023301960222474f,
                             0.47139673682599764856f,
                             0.38268343236508977173f,
                             0.29028467725446236764f,
                             0.19509032201612826785f,
                             0.09801714032956060199f,
                             0.0f};
        // cos(2*pi*e/64), sin(2*pi*e/64) by quadrant from the first-quadrant table
        const int q = (e >> 4) & 3, r = e & 15;
        const float cr = c[r], sr = c[16 - r];
        float cs, sn;
        if (q == 0) {
            cs = cr, sn = sr;
        } else if (q == 1) {
            cs = -sr, sn = cr;
        } else if (q == 2) {
            cs = -cr, sn = -sr;
        } else {
            cs = sr, sn = -cr;
        }
        return cf{cs, -sn};
    }
};

struct Geom64 {
    static constexpr int N = 4096, T = 64, E = 64, SLOTS = 4, CTA = 256;
    static constexpr int ROW = 65;  // padded row of the 64 x 64 exchange frame
    static constexpr size_t SMEM_X = (size_t)SLOTS * 64 * ROW * sizeof(cf);
    static constexpr size_t SMEM_TW = (size_t)64 * 64 * sizeof(cf);
    static constexpr size_t SMEM_W = (size_t)N * sizeof(float);
    static constexpr size_t SMEM = SMEM_X + SMEM_TW + SMEM_W;
};

// phase functions shared with the CPU emulation
template <int IN>
struct SpectrumFrame64 {
    static_assert(IN == FMT_S8 || IN == FMT_U8 || IN == FMT_S16LE, "integer IQ formats only");
    // pass A of column t: points t + 64*n -> DFT64 over n -> times W4096^(c*t) -> exchange row c, column t
    static RFA_HD void pass_a(const char *src /* point t of the frame */, const float *wcol /* &w[t] */,
                              const cf *twcol /* &tw[0][t], row stride 64 */, cf *xcol /* &x[0][t], row stride ROW */) {
        cf u[64];
#pragma unroll
        for (int n = 0; n < 64; n++) {
            const uint32_t raw = (IN == FMT_S16LE) ? ((const uint32_t *)src)[64 * n] : (uint32_t)((const uint16_t *)src)[64 * n];
            u[Dft<64>::in_slot(n)] = decode_point<IN>(raw, wcol[64 * n]);
        }
        Dft<64>::run(u);
        xcol[0] = u[Dft<64>::perm(0)];
#pragma unroll
        for (int c = 1; c < 64; c++) xcol[c * Geom64::ROW] = cmul(u[Dft<64>::perm(c)], twcol[c * 64]);
    }
    // pass B of row k0: DFT64 over the 64 columns -> bins k0 + 64*c
    template <bool PEAK, bool STORE>
    static RFA_HD void pass_b(const cf *xrow /* &x[k0][0] */, float *out /* &row[k0] */, float *pk, float db_bias) {
        cf u[64];
#pragma unroll
        for (int n = 0; n < 64; n++) u[Dft<64>::in_slot(n)] = xrow[n];
        Dft<64>::run(u);
#pragma unroll
        for (int c = 0; c < 64; c++) {
            const float db = logmag_db(u[Dft<64>::perm(c)], db_bias);
            if (STORE) out[(64 * c) ^ 2048] = db;  // fft-shift: bin ^ N/2, k0 < 64 is untouched
            if (PEAK) pk[c] = fmaxf(pk[c], db);
        }
    }
    static RFA_HD int peak_index(int k0, int c) { return (k0 + 64 * c) ^ 2048; }
};

#ifdef __CUDACC__
__device__ __forceinline__ void slot_barrier(int slot) {
    asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
}

// Frames are taken newest first, slot g of G = 4 * gridDim.x takes items g, g + G, ...
// (experiment: static schedule, no time average -- the launcher falls back when avg is requested)
template <int IN>
__global__ void __launch_bounds__(256, 1) spectrum64_kernel(const SpectrumParams p) {
    using G = Geom64;
    using F = SpectrumFrame64<IN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int slot = threadIdx.x >> 6, t = threadIdx.x & 63;
    cf *x = reinterpret_cast<cf *>(smem_raw) + (size_t)slot * 64 * G::ROW;
    cf *tw = reinterpret_cast<cf *>(smem_raw + G::SMEM_X);
    float *w = reinterpret_cast<float *>(smem_raw + G::SMEM_X + G::SMEM_TW);
    for (int i = threadIdx.x; i < 64 * 64; i += G::CTA) tw[i] = p.twN[((i >> 6) * (i & 63)) & 4095];
    for (int i = threadIdx.x; i < G::N; i += G::CTA) w[i] = (p.win ? p.win[i] : 1.0f) * unit_scale<IN>();
    __syncthreads();
    const bool want_peak = p.peaks != nullptr;
    float pk[64];
#pragma unroll
    for (int c = 0; c < 64; c++) pk[c] = -999999.0f;
    constexpr int BPS = in_elem_bytes<IN>();
    const long long nslots = (long long)gridDim.x * G::SLOTS;
    bool worked = false;
    for (long long v = (long long)blockIdx.x * G::SLOTS + slot; v < p.nframes; v += nslots) {
        const long long f = p.nframes - 1 - v;
        F::pass_a((const char *)p.in + (f * (long long)G::N + t) * BPS, w + t, tw + t, x + t);
        slot_barrier(slot);
        float *out = p.rows + frame_row(p, f) * p.row_stride + t;
        if (f >= p.store_from) {
            if (want_peak)
                F::template pass_b<true, true>(x + t * G::ROW, out, pk, p.inv_n2);
            else
                F::template pass_b<false, true>(x + t * G::ROW, out, pk, p.inv_n2);
        } else if (want_peak) {
            F::template pass_b<true, false>(x + t * G::ROW, out, pk, p.inv_n2);
        }
        worked = true;
        slot_barrier(slot);  // the frame buffer is free again
    }
    if (want_peak && worked) {
#pragma unroll
        for (int c = 0; c < 64; c++) atomic_max_float(p.peaks + F::peak_index(t, c), pk[c]);
    }
}
#endif  // __CUDACC__

}  // namespace rfa
