// rfa_fft_core.cuh -- register/shared-memory FFT building blocks for sm_100a.
//
// Replaces the reference's pffft (nativedsp/src/main/cpp/pffft.c:1173-1216 cfftf1_ps,
// :291-425 passf*_ps, :1361-1403 pffft_cplx_finalize, :1324-1359 pffft_zreorder) with a
// Stockham autosort transform: each thread owns 16 (or 32) complex points in registers,
// radix-16/8/4/2 butterflies run entirely in registers and points are exchanged through a
// padded shared-memory frame between passes.  Forward transform, unnormalised,
// X[k] = sum x[n] exp(-2*pi*i*n*k/N), natural-order output -- the contract of
// pffft_transform_ordered(..., PFFFT_FORWARD) (pffft.c:1904).
//
// Everything here is __host__ __device__ so tests/emu can run the exact index logic,
// twiddle addressing and butterflies on the CPU (there is no GPU in the build box).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RFA_HD __host__ __device__ __forceinline__
#define RFA_CX __host__ __device__ constexpr
#else
#define RFA_HD inline
#define RFA_CX constexpr
#endif

namespace rfa {

struct cf {
    float x, y;
};

// Packed FP32 (sm_100a): FADD2 / FMUL2 / FFMA2 operate on an aligned register pair, i.e. on one
// complex value, in ONE issue slot (two FP32-pipe cycles).  SASS operand modifiers give the rest of
// complex arithmetic for free: .LO_HI swaps the halves, .NP/.PN negate one half (multiplication by
// -j / +j) and R.F32 broadcasts a scalar, so a complex add is one instruction and a complex multiply
// two.  ptxas folds the mov.b64 / neg.f32 below into those modifiers (checked with cuobjdump).
// The kernel is issue-bound (profiles/r01_*), so halving the FP32 issue slots is the point;
// tools/ubench/fp32x2.cu measures 0.50 FADD2/clk/SMSP vs 0.98 FADD/clk/SMSP on B200.
#if defined(__CUDA_ARCH__) && !defined(RFA_NO_PACKED)
#define RFA_PACKED 1
typedef unsigned long long rfa_u64;
__device__ __forceinline__ rfa_u64 cpk(float x, float y) {
    rfa_u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
#endif

#ifdef RFA_PACKED
__device__ __forceinline__ cf cunpk(rfa_u64 v) {
    cf a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(v));
    return a;
}
__device__ __forceinline__ cf add2(rfa_u64 a, rfa_u64 b) {
    rfa_u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return cunpk(r);
}
__device__ __forceinline__ cf mul2(rfa_u64 a, rfa_u64 b) {
    rfa_u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return cunpk(r);
}
__device__ __forceinline__ cf fma2(rfa_u64 a, rfa_u64 b, rfa_u64 c) {
    rfa_u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return cunpk(r);
}
#define RFA_PK(v) cpk((v).x, (v).y)
#endif

RFA_HD cf cadd(cf a, cf b) {
#ifdef RFA_PACKED
    return add2(RFA_PK(a), RFA_PK(b));
#else
    return cf{a.x + b.x, a.y + b.y};
#endif
}
RFA_HD cf csub(cf a, cf b) {
#ifdef RFA_PACKED
    return add2(RFA_PK(a), cpk(-b.x, -b.y));
#else
    return cf{a.x - b.x, a.y - b.y};
#endif
}
RFA_HD cf cmul(cf a, cf b) {
#ifdef RFA_EXP_NOBFLY
    return a;
#endif
#ifdef RFA_PACKED
    return fma2(cpk(-a.y, a.x), cpk(b.y, b.y), RFA_PK(mul2(RFA_PK(a), cpk(b.x, b.x))));
#else
    return cf{fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)};
#endif
}
RFA_HD cf mul_mj(cf a) { return cf{a.y, -a.x}; }  // a * (-j)
// a + h*b, a - h*b with a real scalar h (one FFMA2 each)
RFA_HD cf caxpy(float h, cf b, cf a) {
#ifdef RFA_PACKED
    return fma2(RFA_PK(b), cpk(h, h), RFA_PK(a));
#else
    return cf{fmaf(h, b.x, a.x), fmaf(h, b.y, a.y)};
#endif
}
// a * s with a real scalar s
RFA_HD cf cscale(cf a, float s) {
#ifdef RFA_PACKED
    return mul2(RFA_PK(a), cpk(s, s));
#else
    return cf{a.x * s, a.y * s};
#endif
}


// ---------------------------------------------------------------------------
// c2: the same point of TWO frames (A, B), stored as re(A), re(B), im(A), im(B).  With this
// layout every real operation of the transform is one packed instruction over both frames, the
// twiddle / window scalars are shared (R.F32 broadcast) and multiplication by -j is a register
// renaming.  The dual-frame kernel (spectrum2_kernel.cuh) keeps 16 such values per thread.
// ---------------------------------------------------------------------------
struct alignas(16) c2 {
    float rA, rB, iA, iB;
};

RFA_HD c2 cadd(c2 a, c2 b) {
#ifdef RFA_PACKED
    const cf r = add2(cpk(a.rA, a.rB), cpk(b.rA, b.rB)), i = add2(cpk(a.iA, a.iB), cpk(b.iA, b.iB));
    return c2{r.x, r.y, i.x, i.y};
#else
    return c2{a.rA + b.rA, a.rB + b.rB, a.iA + b.iA, a.iB + b.iB};
#endif
}
RFA_HD c2 csub(c2 a, c2 b) {
#ifdef RFA_PACKED
    const cf r = add2(cpk(a.rA, a.rB), cpk(-b.rA, -b.rB)), i = add2(cpk(a.iA, a.iB), cpk(-b.iA, -b.iB));
    return c2{r.x, r.y, i.x, i.y};
#else
    return c2{a.rA - b.rA, a.rB - b.rB, a.iA - b.iA, a.iB - b.iB};
#endif
}
RFA_HD c2 mul_mj(c2 a) { return c2{a.iA, a.iB, -a.rA, -a.rB}; }
RFA_HD c2 cmul(c2 a, cf w) {
#ifdef RFA_EXP_NOBFLY
    return a;
#endif
#ifdef RFA_PACKED
    const rfa_u64 re = cpk(a.rA, a.rB), im = cpk(a.iA, a.iB), wr = cpk(w.x, w.x), wi = cpk(w.y, w.y);
    const cf t = mul2(im, wi), s = mul2(re, wi);
    const cf r = fma2(re, wr, cpk(-t.x, -t.y)), i = fma2(im, wr, RFA_PK(s));
    return c2{r.x, r.y, i.x, i.y};
#else
    return c2{fmaf(a.rA, w.x, -(a.iA * w.y)), fmaf(a.rB, w.x, -(a.iB * w.y)), fmaf(a.iA, w.x, a.rA * w.y),
              fmaf(a.iB, w.x, a.rB * w.y)};
#endif
}
RFA_HD c2 caxpy(float h, c2 b, c2 a) {
#ifdef RFA_PACKED
    const cf r = fma2(cpk(b.rA, b.rB), cpk(h, h), cpk(a.rA, a.rB)), i = fma2(cpk(b.iA, b.iB), cpk(h, h), cpk(a.iA, a.iB));
    return c2{r.x, r.y, i.x, i.y};
#else
    return c2{fmaf(h, b.rA, a.rA), fmaf(h, b.rB, a.rB), fmaf(h, b.iA, a.iA), fmaf(h, b.iB, a.iB)};
#endif
}
RFA_HD c2 cscale(c2 a, float s) {
#ifdef RFA_PACKED
    const cf r = mul2(cpk(a.rA, a.rB), cpk(s, s)), i = mul2(cpk(a.iA, a.iB), cpk(s, s));
    return c2{r.x, r.y, i.x, i.y};
#else
    return c2{a.rA * s, a.rB * s, a.iA * s, a.iB * s};
#endif
}

RFA_CX int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }

// ---------------------------------------------------------------------------
// In-register DFTs.  dftR(u) leaves natural-order output c in u[perm<R>(c)].
// ---------------------------------------------------------------------------
template <class V>
RFA_HD void bfly4(V &a0, V &a1, V &a2, V &a3) {
    V t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mj(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <int R>
struct Dft;

template <>
struct Dft<2> {
    template <class V>
    static RFA_HD void run(V *u) {
        V t = u[0];
        u[0] = cadd(t, u[1]);
        u[1] = csub(t, u[1]);
    }
    static RFA_CX int perm(int c) { return c; }
};

template <>
struct Dft<4> {
    template <class V>
    static RFA_HD void run(V *u) { bfly4(u[0], u[1], u[2], u[3]); }
    static RFA_CX int perm(int c) { return c; }
};

template <>
struct Dft<8> {
    // even part in u[0,2,4,6], odd part in u[1,3,5,7]; X[k] -> u[2k], X[k+4] -> u[2k+1]
    template <class V>
    static RFA_HD void run(V *u) {
        const float h = 0.70710678118654752440f;
        bfly4(u[0], u[2], u[4], u[6]);
        bfly4(u[1], u[3], u[5], u[7]);
        V o1 = cscale(cadd(u[3], mul_mj(u[3])), h);   // * W8^1
        V o2 = mul_mj(u[5]);                          // * W8^2
        V o3 = cscale(csub(mul_mj(u[7]), u[7]), h);   // * W8^3
        V e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6], o0 = u[1];
        u[0] = cadd(e0, o0);
        u[1] = csub(e0, o0);
        u[2] = cadd(e1, o1);
        u[3] = csub(e1, o1);
        u[4] = cadd(e2, o2);
        u[5] = csub(e2, o2);
        u[6] = cadd(e3, o3);
        u[7] = csub(e3, o3);
    }
    static RFA_CX int perm(int c) { return 2 * (c & 3) + (c >> 2); }
};

// radix-4 butterfly whose input a2 is h*s2 (the scaling rides on the adds as FMAs)
template <class V>
RFA_HD void bfly4_h2(V &a0, V &a1, V s2, V &a3, V &o2, float h) {
    V t0 = caxpy(h, s2, a0);
    V t1 = caxpy(-h, s2, a0);
    V t2 = cadd(a1, a3), t3 = mul_mj(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    o2 = csub(t0, t2);
    a3 = csub(t1, t3);
}
// radix-4 butterfly whose inputs a1, a3 are h*s1, h*s3
template <class V>
RFA_HD void bfly4_h13(V &a0, V s1, V &a2, V s3, V &o1, V &o3, float h) {
    V t0 = cadd(a0, a2), t1 = csub(a0, a2);
    V p = cadd(s1, s3), q = mul_mj(csub(s1, s3));
    a0 = caxpy(h, p, t0);
    o1 = caxpy(h, q, t1);
    a2 = caxpy(-h, p, t0);
    o3 = caxpy(-h, q, t1);
}

// Timing experiments only (results are wrong): RFA_EXP_NOBFLY drops the butterflies and twiddle
// products, RFA_EXP_NOXCHG drops the shared-memory exchanges and barriers.
template <>
struct Dft<16> {
    // 4x4 Cooley-Tukey: n = n2 + 4*n1, k = k1 + 4*k2.
    // stage 1: DFT4 over n1 for each n2 -> u[n2 + 4*k1]; twiddle W16^(n2*k1);
    // stage 2: DFT4 over n2 for each k1 -> X[k1 + 4*k2] in u[4*k1 + k2].
    // The four W16^2 / W16^6 twiddles are h*(1-j) / h*(-1-j): the (1-j), (-1-j) parts cost two adds,
    // the factor h = 1/sqrt(2) is folded into the stage-2 additions as FMAs.
    template <class V>
    static RFA_HD void run(V *u) {
#ifdef RFA_EXP_NOBFLY
        return;
#endif
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
        const float h = 0.70710678118654752440f;
#pragma unroll
        for (int n2 = 0; n2 < 4; n2++) bfly4(u[n2], u[n2 + 4], u[n2 + 8], u[n2 + 12]);
        // k1 = 0
        bfly4(u[0], u[1], u[2], u[3]);
        // k1 = 1: W16^1, W16^2, W16^3
        u[5] = cmul(u[5], cf{c1, -s1});
        const V s6 = cadd(u[6], mul_mj(u[6]));  // u6*(1-j)
        u[7] = cmul(u[7], cf{s1, -c1});
        bfly4_h2(u[4], u[5], s6, u[7], u[6], h);
        // k1 = 2: W16^2, W16^4, W16^6
        const V s9 = cadd(u[9], mul_mj(u[9]));      // u9*(1-j)
        u[10] = mul_mj(u[10]);
        const V s11 = csub(mul_mj(u[11]), u[11]);  // u11*(-1-j)
        bfly4_h13(u[8], s9, u[10], s11, u[9], u[11], h);
        // k1 = 3: W16^3, W16^6, W16^9
        u[13] = cmul(u[13], cf{s1, -c1});
        const V s14 = csub(mul_mj(u[14]), u[14]);  // u14*(-1-j)
        u[15] = cmul(u[15], cf{-c1, s1});
        bfly4_h2(u[12], u[13], s14, u[15], u[14], h);
    }
    static RFA_CX int perm(int c) { return 4 * (c & 3) + (c >> 2); }
};

template <>
struct Dft<32> {
    // 2 x 16: DFT16 over the even and the odd inputs, twiddle W32^k1 on the odd half, radix-2 combine.
    // Natural order in, natural order out (everything is registers: the copies are renamings).
    template <class V>
    static RFA_HD void run(V *u) {
        V e[16], o[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            e[i] = u[2 * i];
            o[i] = u[2 * i + 1];
        }
        Dft<16>::run(e);
        Dft<16>::run(o);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int p = Dft<16>::perm(k);
            V t = o[p];
            if (k == 8)
                t = mul_mj(t);
            else if (k > 0)
                t = cmul(t, w32(k));
            u[k] = cadd(e[p], t);
            u[k + 16] = csub(e[p], t);
        }
    }
    static RFA_CX int perm(int c) { return c; }
    static RFA_HD cf w32(int k) {  // exp(-2*pi*i*k/32), 0 < k < 16
        const float c[9] = {1.0f,
                            0.98078528040323044913f,
                            0.92387953251128675613f,
                            0.83146961230254523708f,
                            0.70710678118654752440f,
                            0.55557023301960222474f,
                            0.38268343236508977173f,
                            0.19509032201612826785f,
                            0.0f};
        return k <= 8 ? cf{c[k], -c[8 - k]} : cf{-c[16 - k], -c[k - 8]};
    }
};

// ---------------------------------------------------------------------------
// Frame geometry.  NL = points transformed inside one CTA slot (shared memory),
// T = threads cooperating on the frame, E = NL/T points per thread.
// Radix plan: as many radix-16 passes as fit, then one radix-2/4/8 pass.
// ---------------------------------------------------------------------------
// Radix plan of an n-point transform, shared by the kernels (Plan<NL>) and the host table builder
// (make_pass_twiddles): as many radix-16 passes as fit, then one radix-2/4/8 pass -- except for
// 8192 and 16384 points, where a thread holds 32 points and 16 x 16 x 32 / 16 x 32 x 32 save a whole pass (and exchange).
#ifdef RFA_R8  // timing experiment: N = 4096 as 8 x 8 x 8 x 8 with 8 points per thread (512 threads per frame, <= 64 registers)
RFA_CX int plan_passes(int lg) { return lg == 12 ? 4 : (lg == 14 ? 3 : lg / 4 + (lg % 4 ? 1 : 0)); }
RFA_CX int plan_radix(int lg, int pass) {
    return lg == 12 ? 8 : (lg == 14 ? (pass == 0 ? 16 : 32) : (pass < lg / 4 ? 16 : (1 << (lg % 4))));
}
#elif !defined(RFA_8192_T512)
// N = 8192 is 16 x 16 x 32 with 32 points per thread as well: 256 threads per frame and ONE exchange frame let two CTAs
// share an SM (53.9 against 57.2 us per 2^24 int8 samples for 16 x 16 x 16 x 2 on one 512-thread CTA,
// profiles/r02p_8192_e32_and_average_cta.txt; -DRFA_8192_T512 builds the earlier geometry)
RFA_CX int plan_passes(int lg) { return (lg == 14 || lg == 13) ? 3 : lg / 4 + (lg % 4 ? 1 : 0); }
RFA_CX int plan_radix(int lg, int pass) {
    return lg == 14 ? (pass == 0 ? 16 : 32) : (lg == 13 ? (pass < 2 ? 16 : 32) : (pass < lg / 4 ? 16 : (1 << (lg % 4))));
}
#else
RFA_CX int plan_passes(int lg) { return lg == 14 ? 3 : lg / 4 + (lg % 4 ? 1 : 0); }
RFA_CX int plan_radix(int lg, int pass) {
    return lg == 14 ? (pass == 0 ? 16 : 32) : (pass < lg / 4 ? 16 : (1 << (lg % 4)));
}
#endif

template <int NL>
struct Plan {
    static constexpr int LG = ilog2c(NL);
    static constexpr int PASSES = plan_passes(LG);
    static RFA_CX int radix(int pass) { return plan_radix(LG, pass); }
    // product of the radices of all passes before `pass`
    static RFA_CX int prod(int pass) { return pass == 0 ? 1 : prod(pass - 1) * radix(pass - 1); }
    // padded shared-memory frame: one pad slot per 16 points keeps the radix-R
    // scatter of the first pass (stride R) and the stride-1 gathers conflict free
    static constexpr int SMEM_POINTS = NL + NL / 16;
};

RFA_HD int phys(int a) { return a + (a >> 4); }

// Twiddle tables are stored per pass so that a warp's loads are contiguous: pass with
// radix R and stride product P keeps W_{P*R}^{k*r} at  pass_tw_offset + (r-1)*P + k
// (k = 0..P-1 is the fast index = consecutive lanes).  Total size < NL entries.
template <int NL>
RFA_CX int pass_tw_offset(int pass) {
    return pass <= 1 ? 0 : pass_tw_offset<NL>(pass - 1) + (Plan<NL>::radix(pass - 1) - 1) * Plan<NL>::prod(pass - 1);
}
template <int NL>
RFA_CX int pass_tw_total() {
    return pass_tw_offset<NL>(Plan<NL>::PASSES);
}

// One pass, gather side: butterfly i takes x[i + r*NL/R], r < R, applies the Stockham
// twiddle W_{P*R}^{k*r} (k = i mod P), runs the DFT.  `tw` points at this pass's table.
// `u` holds E/R butterflies of R points each, butterfly b is i = tid + b*T.
// All strides are multiples of 16, so phys(i + r*STR) = phys(i) + r*(STR + STR/16): one
// address per butterfly, the rest are immediate offsets.
template <int NL, int T, int R, int P, class V>
RFA_HD void pass_gather(const V *x, const cf *tw, int tid, V *u) {
    constexpr int E = NL / T;
    constexpr int NB = E / R;
    constexpr int STR = NL / R;
    static_assert(STR % 16 == 0 || NL < 256, "stride must keep the padding pattern");
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int i = tid + b * T;
        const int k = i & (P - 1);
        const cf *twk = tw + k;
        if (STR % 16 == 0) {
            const V *xi = x + phys(i);
#pragma unroll
            for (int r = 0; r < R; r++) {
#ifdef RFA_EXP_NOXCHG
                V v = u[b * R + r];
#else
                V v = xi[r * (STR + STR / 16)];
#endif
                if (P > 1 && r > 0) v = cmul(v, twk[(r - 1) * P]);
                u[b * R + r] = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                V v = x[phys(i + r * STR)];
                if (P > 1 && r > 0) v = cmul(v, twk[(r - 1) * P]);
                u[b * R + r] = v;
            }
        }
        Dft<R>::run(u + b * R);
    }
}

// Twiddle r of a radix-R butterfly (R = 8, 16) composed from the entries a thread KEEPS: t[r-1] for r = 1, 2, 3 and
// r = 4 (8, 12).  W^(k r) = W^(k (r & 3)) * W^(k 4 (r >> 2)): the other entries of a 15- (7-) element register array are
// never read and the compiler drops them.  Used for the last pass's register twiddles: at N = 4096 eighteen registers
// fewer per thread took the kernel from 43.6 to 40.5 us per 2^24 samples (profiles/r02k_composed_last_pass_twiddles.txt).
// KEEP = how many entries of the (R-1)-element array are read: all of them (no composition), six (four for R = 8), or
// two (W^k and W^4k, the other powers by squaring: five more products, ten registers fewer).  Which one wins depends
// on the kernel's register pressure: profiles/r02k_twiddles_kept_sweep.txt.
template <int R, int KEEP = 6>
RFA_HD cf composed_twiddle(const cf *t, int r) {
    static_assert(R == 8 || R == 16, "composition by (r & 3, r >> 2)");
    if (KEEP >= R - 1) return t[r - 1];
    const int lo = r & 3, hi = r >> 2;
    if (KEEP == 2) {
        const cf l1 = t[0], l2 = cmul(l1, l1), l3 = cmul(l2, l1);
        const cf h1 = t[3], h2 = cmul(h1, h1), h3 = cmul(h2, h1);
        const cf L = lo == 1 ? l1 : (lo == 2 ? l2 : l3), H = hi == 1 ? h1 : (hi == 2 ? h2 : h3);
        if (hi == 0) return L;
        if (lo == 0) return H;
        return cmul(L, H);
    }
    if (hi == 0) return t[lo - 1];
    if (lo == 0) return t[4 * hi - 1];
    return cmul(t[lo - 1], t[4 * hi - 1]);
}

// Radix-32 pass with one butterfly per thread whose 31 twiddles W^(k r) are COMPOSED from ten table entries (r = 1, 2, 3
// and r = 4, 8, ... 28) and 21 products: the last pass of the 16 x 32 x 32 plan (N = 16384) has a 127 KB table that
// lives in L2, and 31 eight-byte loads per thread and frame with no registers to prefetch them were 27 % of that
// kernel's stall cycles (ncu long_scoreboard, profiles/r01e_large_n_ncu_summary.txt).  One more rounding per twiddle.
template <int NL, int T, int P, class V>
RFA_HD void pass_gather_r32_composed(const V *x, const cf *tw, int tid, V *u) {
    constexpr int R = 32, STR = NL / R;
    static_assert(NL / T == R && STR % 16 == 0, "one radix-32 butterfly per thread");
    const int k = tid & (P - 1);
    const cf *twk = tw + k;
    cf lo[4], hi[8];
#pragma unroll
    for (int r = 1; r < 4; r++) lo[r] = twk[(r - 1) * P];
#pragma unroll
    for (int j = 1; j < 8; j++) hi[j] = twk[(4 * j - 1) * P];
    const V *xi = x + phys(tid);
#pragma unroll
    for (int r = 0; r < R; r++) {
        V v = xi[r * (STR + STR / 16)];
        if (r > 0) {
            const cf w = (r & 3) == 0 ? hi[r >> 2] : ((r >> 2) == 0 ? lo[r & 3] : cmul(lo[r & 3], hi[r >> 2]));
            v = cmul(v, w);
        }
        u[r] = v;
    }
    Dft<R>::run(u);
}

// The same for a radix-16 pass: 15 twiddles from 6 table entries (r = 1, 2, 3 and 4, 8, 12) and 9 products
template <int NL, int T, int P, class V>
RFA_HD void pass_gather_r16_composed(const V *x, const cf *tw, int tid, V *u) {
    constexpr int R = 16, E = NL / T, NB = E / R, STR = NL / R;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int i = tid + b * T;
        const cf *twk = tw + (i & (P - 1));
        cf lo[4], hi[4];
#pragma unroll
        for (int r = 1; r < 4; r++) lo[r] = twk[(r - 1) * P];
#pragma unroll
        for (int j = 1; j < 4; j++) hi[j] = twk[(4 * j - 1) * P];
        const V *xi = x + phys(i);
#pragma unroll
        for (int r = 0; r < R; r++) {
            V v = xi[r * (STR + STR / 16)];
            if (r > 0) {
                const cf w = (r & 3) == 0 ? hi[r >> 2] : ((r >> 2) == 0 ? lo[r & 3] : cmul(lo[r & 3], hi[r >> 2]));
                v = cmul(v, w);
            }
            u[b * R + r] = v;
        }
        Dft<R>::run(u + b * R);
    }
}

// One pass, scatter side: natural-order output c of butterfly i goes to
// y[(i-k)*R + k + c*P].  For P a multiple of 16 the c-offsets are immediates; the first
// pass (P = 1, R = 16) writes 16 consecutive points, phys(16*i + c) = 17*i + c.
template <int NL, int T, int R, int P, class V>
RFA_HD void pass_scatter(V *y, int tid, const V *u) {
    constexpr int E = NL / T;
    constexpr int NB = E / R;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int i = tid + b * T;
        const int k = i & (P - 1);
        const int j = (i - k) * R + k;
        if (P % 16 == 0) {
            V *yj = y + phys(j);
#pragma unroll
            for (int c = 0; c < R; c++) yj[c * (P + P / 16)] = u[b * R + Dft<R>::perm(c)];
        } else if (P == 1 && R == 16) {
            V *yj = y + 17 * i;
#pragma unroll
            for (int c = 0; c < R; c++) yj[c] = u[b * R + Dft<R>::perm(c)];
        } else {
#pragma unroll
            for (int c = 0; c < R; c++) y[phys(j + c * P)] = u[b * R + Dft<R>::perm(c)];
        }
    }
}

// First pass without the gather: caller filled u[b*R + r] with point (tid + b*T) + r*NL/R.
template <int NL, int T, int R, class V>
RFA_HD void pass_first_compute(V *u) {
    constexpr int NB = (NL / T) / R;
#pragma unroll
    for (int b = 0; b < NB; b++) Dft<R>::run(u + b * R);
}

// Output bin of natural-order output c of butterfly b in the LAST pass (P*R == NL):
// (i-k)*R + k + c*P with k = i  ->  i + c*P.
template <int NL, int T, int R, int P>
RFA_HD int last_pass_bin(int tid, int b, int c) {
    return (tid + b * T) + c * P;
}

// ---------------------------------------------------------------------------
// Sample conversion, bit-exact with the reference's look-up tables.
//   s8 : Signed8BitIQConverter.java:48-50     lut[i] = (i-128)/128.0f, index b+128
//   u8 : Unsigned8BitIQConverter.java:48-50   lut[i] = (i-127.4f)/128.0f
//   s16: Signed16BitIQConverter.kt:46-57      lut[u] = s/32768.0f
// Every value is an integer (or integer minus 127.4f, a multiple of 2^-17 below 128)
// scaled by a power of two, so the arithmetic forms below round nowhere and equal
// the tables for all code points (checked exhaustively in tests).
// ---------------------------------------------------------------------------
enum : int { FMT_S8 = 0, FMT_U8 = 1, FMT_S16LE = 2, FMT_CF32 = 3, FMT_PF32 = 4 };

RFA_HD float conv_s8(int b /* -128..127 */) { return (float)b * 0.0078125f; }
RFA_HD float conv_u8(int b /* 0..255 */) { return ((float)b - 127.4f) * 0.0078125f; }
RFA_HD float conv_s16(int s /* -32768..32767 */) { return (float)s * (1.0f / 32768.0f); }

// The same conversions without an int->float instruction (I2F shares the quarter-rate
// XU pipe with the log2 of the dB stage): the code is dropped into the mantissa of 2^23,
// so  bits(0x4B000000 | u) == 8388608.0f + u  exactly, and one FADD recovers the integer.
// `ws` is the window tap times the power-of-two unit (1/128 or 1/32768): scaling by a power
// of two commutes with rounding, so value*ws == fl(lut[code] * w), the reference's product.
RFA_HD float bits_to_float(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
// PRMT: result byte i = byte (sel >> 4i) & 7 of the pair {a = bytes 0-3, b = bytes 4-7}
RFA_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}
// one PRMT drops byte 0 / byte 1 (or half 0 / half 1) of `raw` into the mantissa of 2^23
RFA_HD float magic_byte0(uint32_t raw) { return bits_to_float(byte_perm(raw, 0x4B000000u, 0x7650u)); }
RFA_HD float magic_byte1(uint32_t raw) { return bits_to_float(byte_perm(raw, 0x4B000000u, 0x7651u)); }
RFA_HD float magic_half0(uint32_t raw) { return bits_to_float(byte_perm(raw, 0x4B000000u, 0x7610u)); }
RFA_HD float magic_half1(uint32_t raw) { return bits_to_float(byte_perm(raw, 0x4B000000u, 0x7632u)); }

// dB scaling of nativedsp.cpp:72-79: 10*log10(sqrt((re/N)^2+(im/N)^2)) = 5*log10(|X|^2/N^2)
//   = 1.50515*log2(|X|^2) - 3.0103*log2(N).  N is a power of two, so the bias is an exact
// multiple of 3.0103 and the division by N^2 never has to be executed.
RFA_HD float logmag_db(cf v, float db_bias) {
    const float pw = fmaf(v.x, v.x, v.y * v.y);
#ifdef __CUDA_ARCH__
    float lg;  // MUFU.LG2 without the denormal pre-scaling: 0 -> -inf like log10f(0)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(pw));
    return fmaf(1.5051499783199060f, lg, db_bias);
#else
    return fmaf(1.5051499783199060f, log2f(pw), db_bias);
#endif
}

// two bins at once: the 1.505*lg + bias scaling as ONE packed FFMA2 (timing experiment -DRFA_DB2)
RFA_HD void logmag_db2(cf v0, cf v1, float db_bias, float &d0, float &d1) {
#if defined(RFA_PACKED) && defined(RFA_DB2)
    const float p0 = fmaf(v0.x, v0.x, v0.y * v0.y), p1 = fmaf(v1.x, v1.x, v1.y * v1.y);
    float l0, l1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(p0));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(p1));
    const cf r = fma2(cpk(l0, l1), cpk(1.5051499783199060f, 1.5051499783199060f), cpk(db_bias, db_bias));
    d0 = r.x;
    d1 = r.y;
#else
    d0 = logmag_db(v0, db_bias);
    d1 = logmag_db(v1, db_bias);
#endif
}

}  // namespace rfa
