// vamp_device.cuh - device-side building blocks shared by every kernel of the hot path.
//
// B200 (sm_100a) design notes
//  * The path is FP32/FP64-FMA + MUFU bound: no tensor cores, almost no HBM
//    traffic.  A *warp* owns one parameter vector ("item" / chain); the 32 lanes own pixels
//    p = j*32 + lane, j < PPL, so spectrum loads are 128-byte coalesced and every per-pixel
//    quantity (tau, flux, residual) lives in registers.  chi^2 / lnL are reduced with xor
//    shuffles, which leaves the bit-identical sum in all lanes (warp-uniform accept decisions).
//  * Pixel coordinates are stored relative to a per-256-pixel tile base (float64 on the host
//    side of the subtraction) so the float32 path never sees the nu ~ 2.4e15 Hz cancellation
//    (SURVEY.md section 7 "hard parts").
//
// Reference arithmetic restated here (vamp_1.0/): GaussFunction vpfits.py:44-54, VoigtFunction
// vpfits.py:58-76 (astropy Voigt1D, wofz form), Tau2flux physics.py:98-105, total vpfits.py:334-338,
// Normal likelihood vpfits.py:39,341, Chisquared vpfits.py:109-118.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define VAMP_FULL_MASK 0xffffffffu
#define VAMP_TILE_PIX 256      // pixels per coordinate tile (= 32 lanes x 8 pixels)
#define VAMP_K2_LONG_PIX 2048  // longest region of the sampler / MAP kernels (long-region objects: 8 tiles in shared memory)
#define VAMP_MAX_PPL 8
#define VAMP_MAX_LSF_TAPS 65

enum VampKind { VAMP_KIND_GAUSS = 0, VAMP_KIND_VOIGT = 1 };

// ------------------------------------------------------------------------------------------------
// Region set (device view).  Built once per upload by vamp_regions_create (vamp_cabi.cu).
struct RegionSetDev {
    int R;
    const int* npix;        // [R]   pixels per region
    const int* pix_off;     // [R]   offset of the region in the pixel arrays (multiple of 64; padding
                            //       pixels hold du = 1e12, flux = 1, 1/noise = 0 and may be read unmasked)
    const int* tile_off;    // [R]   offset of the region's first tile in `ubase`
    const float* du32;      // [Ppad] scaled pixel coordinate minus its tile base
    const float* f32;       // [Ppad] flux
    const float* w32;       // [Ppad] 1/noise (0 in padding)
    const double* du64;
    const double* f64;
    const double* w64;
    // pair layout for the chain kernel: element [jj*32 + lane] holds pixels (64 jj + lane, 64 jj + 32 + lane),
    // i.e. the two rows a warp processes together, so one 8/16-byte load feeds a packed f32x2 operation
    const float2* pdu32;    // [Ppad/2]
    const float2* pf32;
    const float2* pw32;
    const double2* pdu64;
    const double2* pf64;
    const double2* pw64;
    const double* ubase;    // [tiles] scaled coordinate of each tile's base
    const double* lognorm;  // [R]   sum_i ln(1/noise_i) - P/2 ln(2 pi)
    const float* lsf32;     // [taps] or nullptr
    const double* lsf64;
    int lsf_taps;           // odd, 0/1 = identity
    // Multiplets (vamp_regions_set_multiplet): every component of region r carries a SECOND line with tied
    // parameters - in scaled units  amplitude ar A, centroid fr c + sh, Gaussian width fr G (sigma fr s),
    // Lorentzian width lr L.  [R][4] = (fr, sh, ar, lr); ar = 0: the region has single lines; nullptr: none has.
    const double* mult;
};

struct MultD { double fr, sh, ar, lr; };
// The region's multiplet descriptor; false when its components are single lines.
__device__ __forceinline__ bool region_mult(const RegionSetDev& rs, int r, MultD& m) {
    if (!rs.mult) return false;
    m.fr = rs.mult[4 * r + 0];
    m.sh = rs.mult[4 * r + 1];
    m.ar = rs.mult[4 * r + 2];
    m.lr = rs.mult[4 * r + 3];
    return m.ar != 0.0;
}
// The same from the table row (nullptr: single lines) - for callers that must not keep the descriptor live.
__device__ __forceinline__ const double* region_mult_row(const RegionSetDev& rs, int r) {
    if (!rs.mult) return nullptr;
    const double* row = rs.mult + 4 * r;
    return (row[2] != 0.0) ? row : nullptr;
}
// Parameters of line l of an item with n components (l < n: the component itself; n <= l < 2n: the second
// line of component l - n), read from the flat layout theta[4c + i].  K1 / K3 loop over LINES, so a
// multiplet costs them nothing but the longer loop; K2 keeps the sum of both lines in a component's cache slot.
template <int KIND>
__device__ __forceinline__ void line_params(const double* th, int l, int n, const MultD& m, double& a, double& ce,
                                            double& w1, double& w2) {
    const int c = (l < n) ? l : l - n;
    a = th[4 * c + 0];
    ce = th[4 * c + 1];
    w1 = th[4 * c + 2];
    w2 = th[4 * c + 3];
    if (l >= n) {
        a *= m.ar;
        ce = m.fr * ce + m.sh;
        if (KIND == VAMP_KIND_VOIGT) {
            w1 *= m.lr;
            w2 *= m.fr;
        } else {
            w1 *= m.fr;
        }
    }
}

template <typename T> struct RS;
template <> struct RS<float> {
    static __device__ __forceinline__ const float* du(const RegionSetDev& r) { return r.du32; }
    static __device__ __forceinline__ const float* f(const RegionSetDev& r) { return r.f32; }
    static __device__ __forceinline__ const float* w(const RegionSetDev& r) { return r.w32; }
    static __device__ __forceinline__ const float* lsf(const RegionSetDev& r) { return r.lsf32; }
    typedef float2 Pair;
    static __device__ __forceinline__ const float2* pdu(const RegionSetDev& r) { return r.pdu32; }
    static __device__ __forceinline__ const float2* pf(const RegionSetDev& r) { return r.pf32; }
    static __device__ __forceinline__ const float2* pw(const RegionSetDev& r) { return r.pw32; }
};
template <> struct RS<double> {
    static __device__ __forceinline__ const double* du(const RegionSetDev& r) { return r.du64; }
    static __device__ __forceinline__ const double* f(const RegionSetDev& r) { return r.f64; }
    static __device__ __forceinline__ const double* w(const RegionSetDev& r) { return r.w64; }
    static __device__ __forceinline__ const double* lsf(const RegionSetDev& r) { return r.lsf64; }
    typedef double2 Pair;
    static __device__ __forceinline__ const double2* pdu(const RegionSetDev& r) { return r.pdu64; }
    static __device__ __forceinline__ const double2* pf(const RegionSetDev& r) { return r.pf64; }
    static __device__ __forceinline__ const double2* pw(const RegionSetDev& r) { return r.pw64; }
};

// ------------------------------------------------------------------------------------------------
// Scalar helpers per precision.
template <typename T> struct Real;
template <> struct Real<float> {
    // raw MUFU ops (no range fix-up code): 1-2 ulp, denormals flushed - see the accuracy tests
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float ex2(float x) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    // exp(-tau) as exp_scaled(NEGK * tau): the chain kernel keeps tau pre-multiplied by NEGK
    static __device__ __forceinline__ float NEGK() { return -1.4426950408889634f; }
    static __device__ __forceinline__ float exp_scaled(float v) { return ex2(v); }
    static __device__ __forceinline__ float exp_neg(float t) { return ex2(-1.4426950408889634f * t); }
    static __device__ __forceinline__ float exp2_neg(float t) { return ex2(-t); }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
};
template <> struct Real<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double NEGK() { return -1.0; }
    static __device__ __forceinline__ double exp_scaled(double v) { return exp(v); }
    static __device__ __forceinline__ double exp_neg(double t) { return exp(-t); }
    static __device__ __forceinline__ double exp2_neg(double t) { return exp2(-t); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(VAMP_FULL_MASK, v, o);
    return v;   // identical bits in every lane
}

// ------------------------------------------------------------------------------------------------
// Faddeeva function.  H(x, y) = sqrt(pi) * y * Re w(x + i y)  (the unit-amplitude Voigt shape).
//   core  : Weideman (1994) N-term rational, N = 12 (fp32) / 32 (fp64)
//   far   : Laplace continued fraction, 4 (fp32) / 9 (fp64) levels, as a rational in 1/z^2
//   vfar  : pure Lorentzian L^2/(L^2+4d^2) once |z|^2 >= FAR2 (also catches G -> 0 overflow)
// Measured max |H - H_wofz| over the plane: 4.4e-7 (fp32), 7e-15 (fp64)  (tests/test_faddeeva*.py).
// Hand-counted algorithmic flops per evaluation (FMA = 2), used by bench.py's roofline:
//   see vamp_b200/flops.py (fp32: core 71, far 51; fp64: core 151, far 110).
#include "faddeeva_tables.inc"

template <typename T> struct Fad;
template <> struct Fad<float> {
    static constexpr int N = WEID_N_F32;
    static constexpr int CFN = 3, CFD = 3;
    static __device__ __forceinline__ float L() { return WEID_L_F32; }
    static __device__ __forceinline__ float a(int i) { return WEID_A_F32[i]; }
    static __device__ __forceinline__ float cfn(int i) { return CF_NUM_F32[i]; }
    static __device__ __forceinline__ float cfd(int i) { return CF_DEN_F32[i]; }
    // compile-time-index forms: the coefficient becomes an instruction immediate
    static __device__ __forceinline__ constexpr float ac(int i) { return WEID_A_F32_c(i); }
    static __device__ __forceinline__ constexpr float cfnc(int i) { return CF_NUM_F32_c(i); }
    static __device__ __forceinline__ constexpr float cfdc(int i) { return CF_DEN_F32_c(i); }
    static __device__ __forceinline__ float R2() { return 32.0f; }
    static __device__ __forceinline__ float FAR2() { return 1.0e8f; }
    static __device__ __forceinline__ float BIG() { return 1.0e18f; }
};
template <> struct Fad<double> {
    static constexpr int N = WEID_N_F64;
    static constexpr int CFN = 5, CFD = 6;
    static __device__ __forceinline__ double L() { return WEID_L_F64; }
    static __device__ __forceinline__ double a(int i) { return WEID_A_F64[i]; }
    static __device__ __forceinline__ double cfn(int i) { return CF_NUM_F64[i]; }
    static __device__ __forceinline__ double cfd(int i) { return CF_DEN_F64[i]; }
    static __device__ __forceinline__ constexpr double ac(int i) { return WEID_A_F64_c(i); }
    static __device__ __forceinline__ constexpr double cfnc(int i) { return CF_NUM_F64_c(i); }
    static __device__ __forceinline__ constexpr double cfdc(int i) { return CF_DEN_F64_c(i); }
    static __device__ __forceinline__ double R2() { return 64.0; }
    static __device__ __forceinline__ double FAR2() { return 1.0e16; }
    static __device__ __forceinline__ double BIG() { return 1.0e150; }
};

// Mid-fidelity float64 set: Weideman N = 20 (|dH| <= 2.5e-10) and 6 continued-fraction levels beyond |z|^2 = 48
// (<= 4.5e-11).  Rounding-noise free (float64 arithmetic) at 60 % of the cost of the N = 32 set: the second
// stage of K3w's simplex search, whose stopping tolerance (1e-3 in logp) is five orders above its error.
struct FadMid {
    static constexpr int N = WEID_N_MID;
    static constexpr int CFN = 4, CFD = 4;
    static __device__ __forceinline__ double L() { return WEID_L_MID; }
    static __device__ __forceinline__ constexpr double ac(int i) { return WEID_A_MID_c(i); }
    static __device__ __forceinline__ constexpr double cfnc(int i) { return CF_NUM_MID_c(i); }
    static __device__ __forceinline__ constexpr double cfdc(int i) { return CF_DEN_MID_c(i); }
    static __device__ __forceinline__ double R2() { return 48.0; }
};

// H for |z|^2 < R2 (valid over the whole upper half plane): Weideman.  s = x^2 + y^2.
// p(Z) has REAL coefficients, so it is evaluated with the real second-order recurrence
//   b_k = c_k + t b_{k+1} - n b_{k+2},  t = 2 Re Z, n = |Z|^2,   p(Z) = b_1 Z + c_0 - n b_2
// (division of p by (W - Z)(W - conj Z)): 2 FMAs per coefficient instead of the 4 of a complex
// Horner step, and a dependent chain of one FMA per step.  Same accuracy (tests/test_k1_gpu.py).
template <typename T>
__device__ __forceinline__ T fad_core(T x, T y, T s) {
    const T L = Fad<T>::L();
    const T r = Real<T>::rcp(s + (T(2) * L) * y + L * L);        // 1/|L - iz|^2
    const T Zr = (L * L - s) * r;
    const T Zi = (T(2) * L) * x * r;
    const T t = Zr + Zr;
    const T n = Real<T>::fma_(Zr, Zr, Zi * Zi);
    T b1 = Fad<T>::a(0), b2 = T(0);                                // a(j) = c_{N-1-j}
#pragma unroll
    for (int j = 1; j < Fad<T>::N - 1; ++j) {
        const T nb = Real<T>::fma_(t, b1, Real<T>::fma_(-n, b2, Fad<T>::a(j)));
        b2 = b1;
        b1 = nb;
    }
    const T pr = Real<T>::fma_(b1, Zr, Real<T>::fma_(-n, b2, Fad<T>::a(Fad<T>::N - 1)));
    const T pi = b1 * Zi;
    const T ir = (L + y) * r, ii = x * r;                          // 1/(L - iz)
    const T i2r = ir * ir - ii * ii, i2i = T(2) * ir * ii;
    const T q = pr * i2r - pi * i2i;                               // Re(p / (L - iz)^2)
    return y * Real<T>::fma_(T(3.54490770181103205460), q, ir);    // sqrt(pi) y (2 q + ir / sqrt(pi))
}

// H for R2 <= |z|^2 < FAR2 : continued fraction in t = 1/z, v = t^2.
template <typename T>
__device__ __forceinline__ T fad_far(T x, T y, T s) {
    const T rs = Real<T>::rcp(s);
    const T tr = x * rs, ti = -y * rs;
    const T vr = tr * tr - ti * ti, vi = T(2) * tr * ti;
    T nr = Fad<T>::cfn(Fad<T>::CFN - 1), ni = T(0);
#pragma unroll
    for (int k = Fad<T>::CFN - 2; k >= 0; --k) {
        const T a = Real<T>::fma_(nr, vr, Real<T>::fma_(-ni, vi, Fad<T>::cfn(k)));
        ni = Real<T>::fma_(nr, vi, ni * vr);
        nr = a;
    }
    T dr = Fad<T>::cfd(Fad<T>::CFD - 1), di = T(0);
#pragma unroll
    for (int k = Fad<T>::CFD - 2; k >= 0; --k) {
        const T a = Real<T>::fma_(dr, vr, Real<T>::fma_(-di, vi, Fad<T>::cfd(k)));
        di = Real<T>::fma_(dr, vi, di * vr);
        dr = a;
    }
    const T er = nr * dr + ni * di, ei = ni * dr - nr * di;      // N conj(D)
    const T imq = tr * ei + ti * er;                              // Im(t N conj(D))
    return -y * imq * Real<T>::rcp(dr * dr + di * di);
}

// ------------------------------------------------------------------------------------------------
// One absorption component in device units (scaled pixel coordinates of its region).
template <typename T>
struct CompC {
    T A;     // amplitude (tau at line centre for Gauss; amplitude_L for Voigt)
    T cb;    // centroid minus the coordinate-tile base
    T g;     // Gauss: sqrt(log2(e)/2)/sigma ; Voigt: 2 sqrt(ln2)/G
    T y;     // Voigt: sqrt(ln2) L / G
    T y2;    // y*y
    T L2;    // L*L
};

template <typename T, int KIND>
__device__ __forceinline__ CompC<T> make_comp(double A, double c, double w1, double w2, double ubase) {
    CompC<T> cc;
    cc.A = T(A);
    cc.cb = T(c - ubase);
    if (KIND == VAMP_KIND_GAUSS) {
        cc.g = T(0.84932180028801904272 / w1);        // sqrt(0.5*log2(e)) / sigma
        cc.y = cc.y2 = cc.L2 = T(0);
    } else {
        const double big = (double)Fad<T>::BIG();
        cc.g = T(fmin(1.66510922231539551270 / w2, big));       // 2 sqrt(ln 2) / G
        cc.y = T(fmin(0.83255461115769775635 * w1 / w2, big));  // sqrt(ln 2) L / G
        cc.y2 = cc.y * cc.y;
        cc.L2 = T(w1 * w1);
    }
    return cc;
}

// Unit-amplitude shape of one component at PPL pixels of this lane.  All branches are
// warp-uniform (ballots), so no divergence: a j-row of 32 consecutive pixels is almost
// always entirely core or entirely far.
// far-path pixel counter (algorithmic-flop accounting for the roofline; see bench.py): `vm[j]`
// is the ballot of valid pixels of row j, `nfar` accumulates valid pixels that took the
// continued-fraction / Lorentzian path.  COUNT = false compiles it away.
template <typename T, int KIND, int PPL, bool COUNT>
__device__ __forceinline__ void comp_shape_c(const CompC<T>& cc, const T (&du)[PPL], T (&H)[PPL],
                                             const unsigned (&vm)[PPL], unsigned& nfar) {
    if (KIND == VAMP_KIND_GAUSS) {
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const T q = (du[j] - cc.cb) * cc.g;
            H[j] = Real<T>::exp2_neg(q * q);
        }
    } else {
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const T d = du[j] - cc.cb;
            const T x = d * cc.g;
            const T s = Real<T>::fma_(x, x, cc.y2);
            const bool far = !(s < Fad<T>::R2());               // also true for inf / NaN
            const unsigned m = __ballot_sync(VAMP_FULL_MASK, far);
            if (COUNT) nfar += __popc(m & vm[j]);
            T h;
            if (m == 0u) {
                h = fad_core<T>(x, cc.y, s);
            } else {
                const bool vfar = !(s < Fad<T>::FAR2());
                T hf = fad_far<T>(x, cc.y, s);
                if (__any_sync(VAMP_FULL_MASK, vfar)) {
                    const T d2 = T(2) * d;
                    const T lor = cc.L2 * Real<T>::rcp(Real<T>::fma_(d2, d2, cc.L2));
                    hf = vfar ? lor : hf;
                }
                if (m == VAMP_FULL_MASK) {
                    h = hf;
                } else {
                    const T hc = fad_core<T>(x, cc.y, far ? T(0) : s);
                    h = far ? hf : hc;
                }
            }
            H[j] = h;
        }
    }
}

template <typename T, int KIND, int PPL>
__device__ __forceinline__ void comp_shape(const CompC<T>& cc, const T (&du)[PPL], T (&H)[PPL]) {
    unsigned vm[PPL];
#pragma unroll
    for (int j = 0; j < PPL; ++j) vm[j] = 0u;
    unsigned nfar = 0u;
    comp_shape_c<T, KIND, PPL, false>(cc, du, H, vm, nfar);
}

// ------------------------------------------------------------------------------------------------
// Two-wide values.  The chain kernel processes two 32-pixel rows together; in fp32 the pair lives
// in one 64-bit register and every arithmetic step is ONE packed instruction (sm_100
// fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2).  Measured on B200 (tools/micro/ffma2_bench.cu):
// FFMA2 delivers the same 72-74 TFLOP/s as FFMA with half the issue slots, which is what this
// issue-bound kernel needs.  In fp64 the pair is two doubles and the same code compiles to scalar ops.
// Negations are written on the unpacked halves: ptxas folds them into the operand modifiers of the
// consuming FFMA2/FADD2 (an XOR on the 64-bit register would cost two LOP3 each).
template <typename T> struct P2;
template <> struct P2<float> { unsigned long long v; };
template <> struct P2<double> { double a, b; };

__device__ __forceinline__ P2<float> p2_make(float a, float b) {
    P2<float> r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ P2<double> p2_make(double a, double b) { P2<double> r; r.a = a; r.b = b; return r; }
template <typename T> __device__ __forceinline__ P2<T> p2_bc(T a) { return p2_make(a, a); }
__device__ __forceinline__ float p2_lo(P2<float> x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return a; }
__device__ __forceinline__ float p2_hi(P2<float> x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return b; }
__device__ __forceinline__ double p2_lo(P2<double> x) { return x.a; }
__device__ __forceinline__ double p2_hi(P2<double> x) { return x.b; }
__device__ __forceinline__ P2<float> p2_from(float2 v) { return p2_make(v.x, v.y); }
__device__ __forceinline__ P2<double> p2_from(double2 v) { return p2_make(v.x, v.y); }
__device__ __forceinline__ P2<float> p2_fma(P2<float> a, P2<float> b, P2<float> c) {
    P2<float> d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
}
__device__ __forceinline__ P2<float> p2_mul(P2<float> a, P2<float> b) {
    P2<float> d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ P2<float> p2_add(P2<float> a, P2<float> b) {
    P2<float> d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ P2<float> p2_sub(P2<float> a, P2<float> b) {
    P2<float> d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ P2<float> p2_neg(P2<float> a) { return p2_make(-p2_lo(a), -p2_hi(a)); }
__device__ __forceinline__ P2<double> p2_fma(P2<double> a, P2<double> b, P2<double> c) { return p2_make(fma(a.a, b.a, c.a), fma(a.b, b.b, c.b)); }
__device__ __forceinline__ P2<double> p2_mul(P2<double> a, P2<double> b) { return p2_make(a.a * b.a, a.b * b.b); }
__device__ __forceinline__ P2<double> p2_add(P2<double> a, P2<double> b) { return p2_make(a.a + b.a, a.b + b.b); }
__device__ __forceinline__ P2<double> p2_sub(P2<double> a, P2<double> b) { return p2_make(a.a - b.a, a.b - b.b); }
__device__ __forceinline__ P2<double> p2_neg(P2<double> a) { return p2_make(-a.a, -a.b); }
template <typename T> __device__ __forceinline__ P2<T> p2_rcp(P2<T> a) { return p2_make(Real<T>::rcp(p2_lo(a)), Real<T>::rcp(p2_hi(a))); }
template <typename T> __device__ __forceinline__ P2<T> p2_exp_scaled(P2<T> a) { return p2_make(Real<T>::exp_scaled(p2_lo(a)), Real<T>::exp_scaled(p2_hi(a))); }

// fad_core for two rows (the arithmetic of fad_core, every step one packed instruction).
// Z = (L + iz)/(L - iz):  with R = 1/|L - iz|^2 = 1/(s + 2Ly + L^2), i.e. s = 1/R - 2Ly - L^2,
//   Re Z = (L^2 - s) R = 2L(L + y) R - 1,   2 Re Z = 4L(L + y) R - 2,   -|Z|^2 = -(s - 2Ly + L^2) R = 4Ly R - 1 :
// each ONE fused multiply-add of R with a per-component constant (no dependence on s, Re Z or Im Z), and
//   Re(p / (L - iz)^2) = pr (ir^2 - XR^2) - 4L (b1 ir) XR^2     (Im Z = 2L XR never formed).
// 35 packed operations (27 of them FMAs) + 2 MUFU per pixel pair; the form it replaced took 40 (24 FMAs) for the
// same result: max |H - H_wofz| 4.7e-7 against 4.0e-7 in fp32 arithmetic (tests/test_k1_gpu.py::test_faddeeva_plane).
#ifndef VAMP_FAD_CORE35
#define VAMP_FAD_CORE35 1
#endif
template <typename T, typename F = Fad<T>>
__device__ __forceinline__ P2<T> fad_core2(P2<T> X, T y, P2<T> S) {
    const T L = F::L();
#if VAMP_FAD_CORE35
    const T Ly = L + y, K1 = Real<T>::fma_(T(2) * L, y, L * L), K2 = (T(4) * L) * Ly, K3 = (T(4) * L) * y;
    const P2<T> R = p2_rcp<T>(p2_add(S, p2_bc<T>(K1)));
    const P2<T> t = p2_fma(p2_bc<T>(K2), R, p2_bc<T>(T(-2)));
    const P2<T> Zr = p2_fma(p2_bc<T>(T(0.5) * K2), R, p2_bc<T>(T(-1)));
    const P2<T> nn = p2_fma(p2_bc<T>(K3), R, p2_bc<T>(T(-1)));          // -|Z|^2
    const P2<T> XR = p2_mul(X, R);                                      // Im 1/(L - iz)
    const P2<T> ir = p2_mul(p2_bc<T>(Ly), R);                           // Re 1/(L - iz)
    P2<T> b1 = p2_fma(t, p2_bc<T>(F::ac(0)), p2_bc<T>(F::ac(1))), b2 = p2_bc<T>(F::ac(0));
#pragma unroll
    for (int j = 2; j < F::N - 1; ++j) {
        const P2<T> nb = p2_fma(t, b1, p2_fma(nn, b2, p2_bc<T>(F::ac(j))));
        b2 = b1;
        b1 = nb;
    }
    const P2<T> pr = p2_fma(b1, Zr, p2_fma(nn, b2, p2_bc<T>(F::ac(F::N - 1))));
    const P2<T> XR2 = p2_mul(XR, XR);
    const P2<T> i2r = p2_fma(ir, ir, p2_neg(XR2));
    const P2<T> v = p2_mul(p2_mul(b1, ir), XR2);
    const P2<T> q = p2_fma(v, p2_bc<T>(T(-4) * L), p2_mul(pr, i2r));    // Re(p / (L - iz)^2)
    return p2_mul(p2_bc<T>(y), p2_fma(p2_bc<T>(T(3.54490770181103205460)), q, ir));
#else
    const T L2 = L * L, twoLy = (T(2) * L) * y;
    const P2<T> R = p2_rcp<T>(p2_add(S, p2_bc<T>(twoLy + L2)));
    const P2<T> Zr = p2_mul(p2_sub(p2_bc<T>(L2), S), R);
    const P2<T> XR = p2_mul(X, R);                                      // Im 1/(L - iz)
    const P2<T> Zi = p2_mul(XR, p2_bc<T>(T(2) * L));
    const P2<T> t = p2_add(Zr, Zr);
    const P2<T> nn = p2_mul(p2_sub(p2_bc<T>(twoLy - L2), S), R);        // -|Z|^2
    P2<T> b1 = p2_fma(t, p2_bc<T>(F::ac(0)), p2_bc<T>(F::ac(1))), b2 = p2_bc<T>(F::ac(0));
#pragma unroll
    for (int j = 2; j < F::N - 1; ++j) {
        const P2<T> nb = p2_fma(t, b1, p2_fma(nn, b2, p2_bc<T>(F::ac(j))));
        b2 = b1;
        b1 = nb;
    }
    const P2<T> pr = p2_fma(b1, Zr, p2_fma(nn, b2, p2_bc<T>(F::ac(F::N - 1))));
    const P2<T> pi = p2_mul(b1, Zi);
    const P2<T> ir = p2_mul(p2_bc<T>(L + y), R);                        // Re 1/(L - iz)
    const P2<T> i2r = p2_sub(p2_mul(ir, ir), p2_mul(XR, XR));
    const P2<T> i2i = p2_mul(p2_add(ir, ir), XR);
    const P2<T> q = p2_sub(p2_mul(pr, i2r), p2_mul(pi, i2i));           // Re(p / (L - iz)^2)
    return p2_mul(p2_bc<T>(y), p2_fma(p2_bc<T>(T(3.54490770181103205460)), q, ir));
#endif
}

// fad_far for two rows: 28 (fp32: 26 with VAMP_FAD_FAR26) packed operations + 4 MUFU per pixel pair
#ifndef VAMP_FAD_FAR26
#define VAMP_FAD_FAR26 1
#endif
template <typename T, typename F = Fad<T>>
__device__ __forceinline__ P2<T> fad_far2(P2<T> X, T y, P2<T> S) {
    const P2<T> RS_ = p2_rcp<T>(S);
    const P2<T> tr = p2_mul(X, RS_), ti = p2_mul(p2_bc<T>(-y), RS_);
    const P2<T> vr = p2_sub(p2_mul(tr, tr), p2_mul(ti, ti));
    const P2<T> vi = p2_mul(p2_add(tr, tr), ti);
    P2<T> nr, ni, dr, di;
#if VAMP_FAD_FAR26
    if (F::CFN == 3 && F::CFD == 3) {
        // both polynomials are quadratics with REAL coefficients: Re p(v) = p0 + p1 vr + p2 (vr^2 - vi^2),
        // Im p(v) = vi (p1 + 2 p2 vr) - 10 packed operations for the pair instead of the 12 of two complex Horner steps
        const P2<T> a = p2_fma(vr, vr, p2_neg(p2_mul(vi, vi)));
        nr = p2_fma(p2_bc<T>(F::cfnc(2)), a, p2_fma(p2_bc<T>(F::cfnc(1)), vr, p2_bc<T>(F::cfnc(0))));
        ni = p2_mul(vi, p2_fma(p2_bc<T>(T(2) * F::cfnc(2)), vr, p2_bc<T>(F::cfnc(1))));
        dr = p2_fma(p2_bc<T>(F::cfdc(2)), a, p2_fma(p2_bc<T>(F::cfdc(1)), vr, p2_bc<T>(F::cfdc(0))));
        di = p2_mul(vi, p2_fma(p2_bc<T>(T(2) * F::cfdc(2)), vr, p2_bc<T>(F::cfdc(1))));
    } else
#endif
    {
        const P2<T> nvi = p2_neg(vi);
        // Horner in v with real coefficients; the first step has a real accumulator
        nr = p2_fma(p2_bc<T>(F::cfnc(F::CFN - 1)), vr, p2_bc<T>(F::cfnc(F::CFN - 2)));
        ni = p2_mul(p2_bc<T>(F::cfnc(F::CFN - 1)), vi);
#pragma unroll
        for (int k = F::CFN - 3; k >= 0; --k) {
            const P2<T> a = p2_fma(nr, vr, p2_fma(ni, nvi, p2_bc<T>(F::cfnc(k))));
            ni = p2_fma(nr, vi, p2_mul(ni, vr));
            nr = a;
        }
        dr = p2_fma(p2_bc<T>(F::cfdc(F::CFD - 1)), vr, p2_bc<T>(F::cfdc(F::CFD - 2)));
        di = p2_mul(p2_bc<T>(F::cfdc(F::CFD - 1)), vi);
#pragma unroll
        for (int k = F::CFD - 3; k >= 0; --k) {
            const P2<T> a = p2_fma(dr, vr, p2_fma(di, nvi, p2_bc<T>(F::cfdc(k))));
            di = p2_fma(dr, vi, p2_mul(di, vr));
            dr = a;
        }
    }
    const P2<T> er = p2_fma(nr, dr, p2_mul(ni, di));
    const P2<T> ei = p2_sub(p2_mul(ni, dr), p2_mul(nr, di));
    const P2<T> imq = p2_fma(tr, ei, p2_mul(ti, er));
    const P2<T> rd = p2_rcp<T>(p2_fma(dr, dr, p2_mul(di, di)));
    return p2_mul(p2_mul(p2_bc<T>(-y), imq), rd);
}

// Same two rational functions with a different y in each half: K3 evaluates TWO COMPONENTS of one pixel row
// as one two-wide value, which gives its latency-bound float64 objective two independent dependency chains
// (in float64 a P2 operation is simply two adjacent scalar operations).
template <typename T>
__device__ __forceinline__ P2<T> fad_core2y(P2<T> X, P2<T> Y, P2<T> S) {
    const T L = Fad<T>::L();
    const P2<T> L2 = p2_bc<T>(L * L), twoLY = p2_mul(p2_bc<T>(T(2) * L), Y);
    const P2<T> R = p2_rcp<T>(p2_add(S, p2_add(twoLY, L2)));
    const P2<T> Zr = p2_mul(p2_sub(L2, S), R);
    const P2<T> XR = p2_mul(X, R);
    const P2<T> Zi = p2_mul(XR, p2_bc<T>(T(2) * L));
    const P2<T> t = p2_add(Zr, Zr);
    const P2<T> nn = p2_mul(p2_sub(p2_sub(twoLY, L2), S), R);
    P2<T> b1 = p2_fma(t, p2_bc<T>(Fad<T>::ac(0)), p2_bc<T>(Fad<T>::ac(1))), b2 = p2_bc<T>(Fad<T>::ac(0));
#pragma unroll
    for (int j = 2; j < Fad<T>::N - 1; ++j) {
        const P2<T> nb = p2_fma(t, b1, p2_fma(nn, b2, p2_bc<T>(Fad<T>::ac(j))));
        b2 = b1;
        b1 = nb;
    }
    const P2<T> pr = p2_fma(b1, Zr, p2_fma(nn, b2, p2_bc<T>(Fad<T>::ac(Fad<T>::N - 1))));
    const P2<T> pi = p2_mul(b1, Zi);
    const P2<T> ir = p2_mul(p2_add(p2_bc<T>(L), Y), R);
    const P2<T> i2r = p2_sub(p2_mul(ir, ir), p2_mul(XR, XR));
    const P2<T> i2i = p2_mul(p2_add(ir, ir), XR);
    const P2<T> q = p2_sub(p2_mul(pr, i2r), p2_mul(pi, i2i));
    return p2_mul(Y, p2_fma(p2_bc<T>(T(3.54490770181103205460)), q, ir));
}

template <typename T>
__device__ __forceinline__ P2<T> fad_far2y(P2<T> X, P2<T> Y, P2<T> S) {
    const P2<T> RS_ = p2_rcp<T>(S);
    const P2<T> mY = p2_neg(Y);
    const P2<T> tr = p2_mul(X, RS_), ti = p2_mul(mY, RS_);
    const P2<T> vr = p2_sub(p2_mul(tr, tr), p2_mul(ti, ti));
    const P2<T> vi = p2_mul(p2_add(tr, tr), ti);
    const P2<T> nvi = p2_neg(vi);
    P2<T> nr = p2_fma(p2_bc<T>(Fad<T>::cfnc(Fad<T>::CFN - 1)), vr, p2_bc<T>(Fad<T>::cfnc(Fad<T>::CFN - 2)));
    P2<T> ni = p2_mul(p2_bc<T>(Fad<T>::cfnc(Fad<T>::CFN - 1)), vi);
#pragma unroll
    for (int k = Fad<T>::CFN - 3; k >= 0; --k) {
        const P2<T> a = p2_fma(nr, vr, p2_fma(ni, nvi, p2_bc<T>(Fad<T>::cfnc(k))));
        ni = p2_fma(nr, vi, p2_mul(ni, vr));
        nr = a;
    }
    P2<T> dr = p2_fma(p2_bc<T>(Fad<T>::cfdc(Fad<T>::CFD - 1)), vr, p2_bc<T>(Fad<T>::cfdc(Fad<T>::CFD - 2)));
    P2<T> di = p2_mul(p2_bc<T>(Fad<T>::cfdc(Fad<T>::CFD - 1)), vi);
#pragma unroll
    for (int k = Fad<T>::CFD - 3; k >= 0; --k) {
        const P2<T> a = p2_fma(dr, vr, p2_fma(di, nvi, p2_bc<T>(Fad<T>::cfdc(k))));
        di = p2_fma(dr, vi, p2_mul(di, vr));
        dr = a;
    }
    const P2<T> er = p2_fma(nr, dr, p2_mul(ni, di));
    const P2<T> ei = p2_sub(p2_mul(ni, dr), p2_mul(nr, di));
    const P2<T> imq = p2_fma(tr, ei, p2_mul(ti, er));
    const P2<T> rd = p2_rcp<T>(p2_fma(dr, dr, p2_mul(di, di)));
    return p2_mul(p2_mul(mY, imq), rd);
}

// ------------------------------------------------------------------------------------------------
// Row-streaming evaluation (K2 / K3): one 32-pixel row per call, path chosen per ROW.
// component constants with the divisions done in T (double divisions cost ~20 instructions each
// and are executed redundantly by every lane once per scalar step).  g and y are clamped so
// that the padding coordinate (1e12) never overflows x^2 + y^2: a Gaussian FWHM below
// 1.7e-6 px (fp32) is a delta function between pixels either way.
// 1/G is clamped (g and y together, so y/x = L/(2 d) - the Lorentzian limit - is preserved):
//   K2 (regions <= 256 px, padding coordinate 1e12):  1/G <= 6e5   ->  x <= 1e18, x^2 + y^2 finite in fp32
//   K1 (regions <= 2^17 px, no padding reads):          1/G <= 6e13  ->  x <= 1.3e19
// A Gaussian FWHM below the clamp is a delta function between pixels either way.
template <typename T> struct PadClamp;
template <> struct PadClamp<float> {
    static __device__ __forceinline__ float ig() { return 6.0e5f; }
    static __device__ __forceinline__ float ig_k1() { return 6.0e13f; }
};
template <> struct PadClamp<double> {
    static __device__ __forceinline__ double ig() { return 1.0e100; }
    static __device__ __forceinline__ double ig_k1() { return 1.0e100; }
};

template <typename T, int KIND>
__device__ __forceinline__ CompC<T> make_comp_fast(T A, T cb, T w1, T w2) {
    CompC<T> cc;
    cc.A = A;
    cc.cb = cb;
    if (KIND == VAMP_KIND_GAUSS) {
        cc.g = T(0.84932180028801904272) * fmin(Real<T>::rcp(w1), PadClamp<T>::ig());
        cc.y = cc.y2 = cc.L2 = T(0);
    } else {
        const T ig = fmin(Real<T>::rcp(w2), PadClamp<T>::ig());
        cc.g = T(1.66510922231539551270) * ig;
        cc.y = T(0.83255461115769775635) * w1 * ig;
        cc.y2 = cc.y * cc.y;
        cc.L2 = w1 * w1;
    }
    return cc;
}

// K1's component constants: divisions in double (one per item and component), same clamp policy.
// PADDED: the caller reads the padded pair-layout arrays (padding coordinate 1e12): K2's clamp.
template <typename T, int KIND, bool PADDED = false>
__device__ __forceinline__ CompC<T> make_comp_k1(double A, double c, double w1, double w2, double ubase) {
    CompC<T> cc;
    cc.A = T(A);
    cc.cb = T(c - ubase);
    const double igmax = PADDED ? (double)PadClamp<T>::ig() : (double)PadClamp<T>::ig_k1();
    if (KIND == VAMP_KIND_GAUSS) {
        cc.g = T(0.84932180028801904272 * fmin(1.0 / w1, igmax));
        cc.y = cc.y2 = cc.L2 = T(0);
    } else {
        const double ig = fmin(1.0 / w2, igmax);
        cc.g = T(1.66510922231539551270 * ig);
        cc.y = T(0.83255461115769775635 * w1 * ig);
        cc.y2 = cc.y * cc.y;
        cc.L2 = T(w1 * w1);
    }
    return cc;
}

// shape of one row (32 pixels); every branch is warp-uniform and each body exists once.
// Padding pixels (du = 1e12) take the Lorentzian limit and evaluate to exactly 0.
template <typename T, int KIND>
__device__ __forceinline__ T row_shape(const CompC<T>& cc, T du, unsigned& nfar) {
    if (KIND == VAMP_KIND_GAUSS) {
        const T q = (du - cc.cb) * cc.g;
        return Real<T>::exp2_neg(q * q);
    }
    const T d = du - cc.cb;
    const T x = d * cc.g;
    const T s = Real<T>::fma_(x, x, cc.y2);
    const bool far = !(s < Fad<T>::R2());
    const unsigned m = __ballot_sync(VAMP_FULL_MASK, far);
    const bool vfar = !(s < Fad<T>::FAR2());
    if (__any_sync(VAMP_FULL_MASK, vfar)) {
        // overflow guard (padding pixels, G -> 0): per-lane select of all three forms
        nfar += __popc(m);
        const T d2 = T(2) * d;
        const T lor = cc.L2 * Real<T>::rcp(Real<T>::fma_(d2, d2, cc.L2));
        const T hf = fad_far<T>(x, cc.y, vfar ? Fad<T>::R2() : s);
        const T hc = fad_core<T>(far ? T(0) : x, cc.y, far ? T(0) : s);
        return vfar ? lor : (far ? hf : hc);
    }
    if (m == VAMP_FULL_MASK) {
        nfar += 32u;
        return fad_far<T>(x, cc.y, s);
    }
    // mixed rows: the Weideman form is valid over the whole plane, so the row takes it as a whole
    return fad_core<T>(x, cc.y, s);
}

// Two rows at once (the chain kernel's unit of work).  Pair policy: all 64 pixels far -> continued
// fraction, otherwise Weideman (valid over the whole upper half plane) for the pair.  No overflow
// guard: make_comp_fast clamps g and y, so padding pixels give H ~ 1e-18 (exp(-tau) = 1 exactly).
// Split in two so the chain kernel can classify pair j + 1 (load, 3 packed operations, vote)
// under the arithmetic of pair j.
template <typename T, int KIND, typename F = Fad<T>>
__device__ __forceinline__ void pair_classify(const CompC<T>& cc, P2<T> DU, P2<T>& X, P2<T>& S, bool& far) {
    X = p2_mul(p2_sub(DU, p2_bc<T>(cc.cb)), p2_bc<T>(cc.g));
    if (KIND == VAMP_KIND_GAUSS) {
        S = p2_mul(X, X);
        far = false;
    } else {
        S = p2_fma(X, X, p2_bc<T>(cc.y2));
        far = __all_sync(VAMP_FULL_MASK, !(fmin(p2_lo(S), p2_hi(S)) < F::R2()));
    }
}
template <typename T, int KIND, typename F = Fad<T>>
__device__ __forceinline__ P2<T> pair_eval(const CompC<T>& cc, P2<T> X, P2<T> S, bool far, unsigned& nfar) {
    if (KIND == VAMP_KIND_GAUSS) return p2_make(Real<T>::exp2_neg(p2_lo(S)), Real<T>::exp2_neg(p2_hi(S)));
    if (far) {
        nfar += 64u;
        return fad_far2<T, F>(X, cc.y, S);
    }
    return fad_core2<T, F>(X, cc.y, S);
}
template <typename T, int KIND>
__device__ __forceinline__ P2<T> pair_shape(const CompC<T>& cc, P2<T> DU, bool& far) {
    P2<T> X, S;
    unsigned dummy = 0u;
    pair_classify<T, KIND>(cc, DU, X, S, far);
    return pair_eval<T, KIND>(cc, X, S, far, dummy);
}

// ------------------------------------------------------------------------------------------------
// Per-lane view of one region tile (<= 32*PPL pixels).
template <typename T, int PPL>
struct LaneTile {
    T du[PPL];
    T f[PPL];
    T w[PPL];          // 1/noise; 0 for p >= P
    unsigned valid;    // bit j set when pixel j*32+lane < P
};

template <typename T, int PPL>
__device__ __forceinline__ void load_tile(LaneTile<T, PPL>& t, const RegionSetDev& rs, int off, int P, int lane) {
    t.valid = 0u;
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        const int p = j * 32 + lane;
        const bool ok = p < P;
        // padding inside the 32-pixel granule is initialised (du far away, f = 1, w = 0) but rows
        // beyond the granule are not: guard the load.
        t.du[j] = ok ? RS<T>::du(rs)[off + p] : T(0);
        t.f[j] = ok ? RS<T>::f(rs)[off + p] : T(1);
        t.w[j] = ok ? RS<T>::w(rs)[off + p] : T(0);
        t.valid |= ok ? (1u << j) : 0u;
    }
}

// LSF: out[p] = 1 + sum_t lsf[t] * (m[p + h - t] - 1), zero outside [0, P).  `buf` is a
// warp-private shared buffer of 32*PPL + 2h elements, `lsf_s` the taps in shared memory.
template <typename T, int PPL>
__device__ __forceinline__ void lsf_apply(T (&m)[PPL], unsigned valid, T* buf, const T* lsf_s, int taps, int lane) {
    const int h = taps >> 1;
    __syncwarp();
    for (int i = lane; i < 2 * h; i += 32) {           // zero both halos
        const int idx = (i < h) ? i : (32 * PPL + i);
        buf[idx] = T(0);
    }
#pragma unroll
    for (int j = 0; j < PPL; ++j) buf[h + j * 32 + lane] = ((valid >> j) & 1u) ? (m[j] - T(1)) : T(0);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        T acc = T(1);
        const T* b = buf + j * 32 + lane + 2 * h;
        for (int t = 0; t < taps; ++t) acc = Real<T>::fma_(lsf_s[t], b[-t], acc);
        m[j] = acc;
    }
}

// Sum of squared residuals of this lane: S = sum (f-m)^2 over valid pixels, X = sum ((f-m) w)^2.
template <typename T, int PPL>
__device__ __forceinline__ void lane_residuals(const LaneTile<T, PPL>& t, const T (&m)[PPL], T& S, T& X) {
    S = T(0);
    X = T(0);
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        const T r = ((t.valid >> j) & 1u) ? (t.f[j] - m[j]) : T(0);
        S = Real<T>::fma_(r, r, S);
        const T rw = r * t.w[j];
        X = Real<T>::fma_(rw, rw, X);
    }
}

// Per-translation-unit instantiation: every kernel file is compiled once per (precision, profile
// kind) with -DVAMP_INST_F64=0|1 -DVAMP_INST_VOIGT=0|1 so the 12 objects build in parallel.
#if defined(VAMP_INST_F64) && defined(VAMP_INST_VOIGT)
#if VAMP_INST_F64
typedef double InstT;
#define VAMP_TN f64
#else
typedef float InstT;
#define VAMP_TN f32
#endif
#if VAMP_INST_VOIGT
#define VAMP_KN voigt
#else
#define VAMP_KN gauss
#endif
constexpr int InstKind = VAMP_INST_VOIGT ? VAMP_KIND_VOIGT : VAMP_KIND_GAUSS;
#define VAMP_CAT3_(a, b, c) a##_##b##_##c
#define VAMP_CAT3(a, b, c) VAMP_CAT3_(a, b, c)
#define VAMP_INST_NAME(name) VAMP_CAT3(name, VAMP_TN, VAMP_KN)
#endif

// Parameter vector layout shared by host and device (scaled units):
//   theta[4*c + 0..3] = (amplitude, centroid, L_fwhm | sigma, G_fwhm | unused),  c < ncomp
//   theta[4*cmax]     = sd   (only when the noise parameter is free)
#define VAMP_TH(c, i) (4 * (c) + (i))

// Full model evaluation for a warp-owned item: tau summed over components in registers,
// m = exp(-tau), optional LSF.  Leaves the model flux in m[].
// Rows are processed two at a time in the packed two-row form (pair_classify / pair_eval: one
// FFMA2 per step for both rows, path chosen per 64-pixel pair); a single-row tile (<= 32 px)
// takes the scalar form.
// MULT (a separate kernel instantiation, launched for region sets with a multiplet table - the single-line
// kernel keeps its registers): md != nullptr marks a multiplet region.
template <typename T, int KIND, int PPL, bool MULT = false>
__device__ __forceinline__ void eval_model(const LaneTile<T, PPL>& t, const double* __restrict__ theta, int ncomp,
                                           double ubase, T (&m)[PPL], T* lsf_buf, const T* lsf_s, int taps, int lane,
                                           const MultD* md = nullptr) {
    // lines to evaluate: the components, then (multiplet regions) their second lines
    const int nlines = (MULT && md) ? 2 * ncomp : ncomp;
    MultD mz;
    mz.fr = mz.sh = mz.ar = mz.lr = 0.0;
    const MultD& mm = (MULT && md) ? *md : mz;
    if (PPL == 1) {
        T tau[PPL];
#pragma unroll
        for (int j = 0; j < PPL; ++j) tau[j] = T(0);
        for (int c = 0; c < nlines; ++c) {
            double pa, pce, pw1, pw2;
            line_params<KIND>(theta, c, ncomp, mm, pa, pce, pw1, pw2);
            const CompC<T> cc = make_comp<T, KIND>(pa, pce, pw1, pw2, ubase);
            T H[PPL];
            comp_shape<T, KIND, PPL>(cc, t.du, H);
#pragma unroll
            for (int j = 0; j < PPL; ++j) tau[j] = Real<T>::fma_(cc.A, H[j], tau[j]);
        }
#pragma unroll
        for (int j = 0; j < PPL; ++j) m[j] = Real<T>::exp_neg(tau[j]);
    } else {
        constexpr int NPR = (PPL + 1) / 2;
        P2<T> tau2[NPR];
#pragma unroll
        for (int jj = 0; jj < NPR; ++jj) tau2[jj] = p2_bc<T>(T(0));
        unsigned nfar = 0u;
        for (int c0 = 0; c0 < nlines; c0 += 32) {
            // lane l forms the constants of line c0 + l (the float64 divisions run once per line, in
            // parallel, instead of nlines times per lane) and they are handed out with shuffles
            CompC<T> mine;
            mine.A = mine.cb = mine.g = mine.y = mine.y2 = mine.L2 = T(0);
            if (c0 + lane < nlines) {
                double pa, pce, pw1, pw2;
                line_params<KIND>(theta, c0 + lane, ncomp, mm, pa, pce, pw1, pw2);
                mine = make_comp_k1<T, KIND>(pa, pce, pw1, pw2, ubase);
            }
            const int cend = (nlines - c0 < 32) ? nlines - c0 : 32;
            for (int l = 0; l < cend; ++l) {
            CompC<T> cc;
            cc.A = __shfl_sync(VAMP_FULL_MASK, mine.A, l);
            cc.cb = __shfl_sync(VAMP_FULL_MASK, mine.cb, l);
            cc.g = __shfl_sync(VAMP_FULL_MASK, mine.g, l);
            cc.y = __shfl_sync(VAMP_FULL_MASK, mine.y, l);
            cc.y2 = cc.y * cc.y;
            cc.L2 = T(0);
            const P2<T> A2 = p2_bc<T>(Real<T>::NEGK() * cc.A);
#pragma unroll
            for (int jj = 0; jj < NPR; ++jj) {
                P2<T> X, S;
                bool far;
                pair_classify<T, KIND>(cc, p2_make(t.du[2 * jj], t.du[(2 * jj + 1 < PPL) ? 2 * jj + 1 : 2 * jj]), X, S, far);
                tau2[jj] = p2_fma(A2, pair_eval<T, KIND>(cc, X, S, far, nfar), tau2[jj]);
            }
            }
        }
#pragma unroll
        for (int jj = 0; jj < NPR; ++jj) {
            const P2<T> M = p2_exp_scaled<T>(tau2[jj]);
            m[2 * jj] = p2_lo(M);
            if (2 * jj + 1 < PPL) m[2 * jj + 1] = p2_hi(M);
        }
    }
    if (taps > 1) lsf_apply<T, PPL>(m, t.valid, lsf_buf, lsf_s, taps, lane);
}

// Gaussian log-likelihood from the reduced sums (double, warp-uniform).
//   sd free (vpfits.py:39,341):      lnL = -S/(2 sd^2) - P ln sd - P/2 ln 2pi
//   per-pixel noise (fit.py:127-171): lnL = -X/2 + sum ln(1/noise) - P/2 ln 2pi
__device__ __forceinline__ double lnL_sd(double S, double sd, int P) {
    return -0.5 * S / (sd * sd) - P * (log(sd) + 0.91893853320467274178);
}
__device__ __forceinline__ double lnL_noise(double X, double lognorm) { return -0.5 * X + lognorm; }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: chains are keyed by a global chain id so
// results do not depend on how regions are sharded over GPUs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * 2.3283064365386962890625e-10; }  // (0,1)
