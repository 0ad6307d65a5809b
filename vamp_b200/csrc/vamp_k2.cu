// vamp_k2.cu - K2: batched Metropolis sampler, one warp per chain, whole chains on device.
//
// Replaces the reference's hot loop  mcmc.sample(iter, burn, thin)  (vamp_1.0/vpfits.py:393,
// 423-425): PyMC 2's auto-assigned scalar Metropolis steppers with adaptive-scale tuning
// (restated in oracle/sampler.py).  Thousands of independent chains - every candidate
// component count of the BIC sweep and every restart - advance inside ONE launch; there is no
// host round trip, no kernel-per-iteration, and no HBM traffic beyond the initial spectrum
// load and the optional tallies.
//
// Per chain (= warp), 4 chains per CTA:
//   * the region's pixels (du, flux, [1/noise]) are staged once per CTA in shared memory, in pair
//     layout, and the lanes sweep them two row pairs (128 pixels) per trip with packed f32x2 math;
//   * the unit-amplitude shape H_c of every component is cached in warp-private shared
//     memory, so a single-parameter step re-evaluates ONE component (P Voigt evaluations) -
//     the same economy PyMC 2's LazyFunction cache gives the reference - and an amplitude or
//     noise-parameter step re-evaluates none;
//   * sum (f-m)^2 is reduced with xor shuffles: every lane holds the identical value and takes
//     the identical accept decision (no divergence, no broadcast);
//   * Philox4x32-10 keyed by (seed, global chain id): lane l draws the Box-Muller pair and the
//     two log-uniforms of steps l and 32 + l of a 64-step block, handed out with one shuffle each;
//   * tallies (i >= burn and i % thin == 0) feed Welford means / co-moments (the reference's
//     mcmc.stats() and chain_covariance, vpfits.py:432-456) and an optional trace.
// DESIGN.md section 4 has the measured history of this kernel and what bounds it.
#include "vamp_device.cuh"
#include "vamp_kernels.h"

namespace {

#ifndef VAMP_K2_TRIPUNROLL
#define VAMP_K2_TRIPUNROLL 1      // unroll factor of the trip loop of a pass (1: rolled - the step loop must fit the instruction cache)
#endif
#ifndef VAMP_K2_WARPS
#define VAMP_K2_WARPS 4           // chains (warps) per CTA; with VAMP_K2_MINBLOCKS CTAs per SM (96 registers: 4 x 5 = 20 warps)
#endif
constexpr int K2_WARPS = VAMP_K2_WARPS;
constexpr int K2_TRIPUNROLL = VAMP_K2_TRIPUNROLL;
#ifndef VAMP_INST_MULT
#define VAMP_INST_MULT 0          // 1: this object holds the multiplet (two tied lines per component) instantiations
#endif
#ifndef VAMP_K2_MINBLOCKS
#define VAMP_K2_MINBLOCKS 5
#endif
#ifndef VAMP_K2_MINBLOCKS64
#define VAMP_K2_MINBLOCKS64 2     // float64 objects: CTAs per SM the register allocation aims at (190 registers).  3 (158 registers, no
#endif                            //    spills): 11 % faster on a 600-iteration launch, 2 % SLOWER on the bench's 3000-iteration one (measured)
#ifndef VAMP_K2_STEPUNROLL
#define VAMP_K2_STEPUNROLL 1      // 1: the centroid / width steps of a component are unrolled (no selects)
#endif
#ifndef VAMP_K2_DRAWAHEAD
#define VAMP_K2_DRAWAHEAD 0       // 1: a step's random pair is shuffled out one step ahead (under the previous pass)
#endif
#ifndef VAMP_K2_SMEMSUM
#define VAMP_K2_SMEMSUM 0         // 1 (fp32): the warp sum goes through shared memory (one store, 8 vector loads, a packed
#endif                            //    adder tree in every lane) instead of five dependent shuffle + add rounds
#ifndef VAMP_K2_NOCOUNT
#define VAMP_K2_NOCOUNT 0         // 1 (experiment only): no work counters - what the accounting of bench.py's numerator costs
#endif
#ifndef VAMP_K2_SMEMRNG
#define VAMP_K2_SMEMRNG 1         // 1: the 64 (normal, log-uniform) pairs of a Philox block are parked in shared memory; a step reads its
#endif                            //    pair with one broadcast load instead of two shuffles (same stream, same values)
#ifndef VAMP_K2_FUSEBASE
#define VAMP_K2_FUSEBASE 1        // 1: the move of `base` to a component runs inside its amplitude pass (FuseArgs)
#endif
constexpr double FWHM_PER_SIGMA = 2.35482004503094938202;

// ------------------------------------------------------------------------------------------------
// Per-chain random stream.  One Philox4x32-10 call per lane yields the (normal, log-uniform)
// pairs of TWO steps of a 64-step block: lane l owns steps l (z0, lu0) and 32 + l (z1, lu1), so
// the hand-out is a plain shuffle from lane (step & 31) of the current half's registers.
// Kept in the kernel's precision (the fp32 kernel runs Box-Muller on the MUFU units).
template <typename T> struct RngMath;
template <> struct RngMath<double> {
    static __device__ __forceinline__ void fill(const uint4& x, double& z0, double& z1, double& l0, double& l1) {
        const double rad = sqrt(-2.0 * log(u01(x.x)));
        double sn, cs;
        sincospi(2.0 * u01(x.y), &sn, &cs);
        z0 = rad * cs;
        z1 = rad * sn;
        l0 = log(u01(x.z));
        l1 = log(u01(x.w));
    }
};
template <> struct RngMath<float> {
    static __device__ __forceinline__ float u(uint32_t x) { return fmaf((float)x, 2.3283064365386962890625e-10f, 1.16415321826934814453125e-10f); }
    static __device__ __forceinline__ void fill(const uint4& x, float& z0, float& z1, float& l0, float& l1) {
        const float rad = sqrtf(-2.0f * __logf(u(x.x)));
        float sn, cs;
        __sincosf(6.283185307179586f * u(x.y), &sn, &cs);
        z0 = rad * cs;
        z1 = rad * sn;
        l0 = __logf(u(x.z));
        l1 = __logf(u(x.w));
    }
};

// One refill of the 64-step block (executed once per 32 steps).  VAMP_K2_RNG_CALL=1 keeps ONE out-of-line copy of
// Philox + Box-Muller instead of one inlined copy per draw site (five in mcmc_chain, ~100 instructions each): the
// step loop then has 8 KB less code between its hot parts (32 KB instruction cache).
#ifndef VAMP_K2_RNG_CALL
#define VAMP_K2_RNG_CALL 1
#endif
template <typename T>
__device__ __noinline__ void rng_refill_call(unsigned long long uid, unsigned long long qblk, unsigned lane, unsigned long long seed,
                                             T* out4) {
    const uint4 x = philox4x32_10(make_uint4((uint32_t)uid, (uint32_t)(uid >> 32), (uint32_t)qblk,
                                             ((uint32_t)(qblk >> 32) << 8) | (uint32_t)lane),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    RngMath<T>::fill(x, out4[0], out4[1], out4[2], out4[3]);
}

// PyMC 2.3 Metropolis.tune switch (oracle/sampler.py:pymc2_tune)
__device__ __forceinline__ float tune_factor(float rate) {
    if (rate < 0.001f) return 0.1f;
    if (rate < 0.05f) return 0.5f;
    if (rate < 0.2f) return 0.9f;
    if (rate > 0.95f) return 10.0f;
    if (rate > 0.75f) return 2.0f;
    if (rate > 0.5f) return 1.1f;
    return 1.0f;
}

// natural log in the step logic: the fp32 kernel uses the raw MUFU lg2 (arguments are O(1))
template <typename T> struct StepMath;
template <> struct StepMath<float> {
    static __device__ __forceinline__ float log_(float x) {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r * 0.693147180559945309f;
    }
};
template <> struct StepMath<double> {
    static __device__ __forceinline__ double log_(double x) { return log(x); }
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ------------------------------------------------------------------------------------------------
// Shared memory.  Pixel buffers are in PAIR layout (element 32 jj + lane holds pixels 64 jj + lane
// and 64 jj + 32 + lane) and sized for one 256-pixel tile, so every buffer sits at a compile-time
// offset from a base register.
//   region slice, shared by the warps of a CTA (they almost always fit the same region, see
//     k2_mcmc_kernel):  du | flux | [1/noise]
//   chain slice (one per warp):  base (tau of the other components, pre-multiplied by -log2 e) |
//     shape cache: ncomp_max + 1 slots (one spare: a proposal is evaluated into the spare slot and
//     accepting it swaps two slot indices instead of copying) | [LSF staging] |
//     sd record | one record per component
template <typename T>
struct alignas(16) CompRec {       // 64 bytes (fp32) / 80 bytes (fp64)
    T th[4];                       // amplitude, centroid, w1, w2 (kernel precision)
    float scl[4];                  // proposal scale asf * psd
    int acc[4];                    // accepted since the last tune
    int slot;                      // shape-cache slot of this component
    float g, y;                    // make_comp_fast constants of (w1, w2) - a centroid step reuses them
    int pad;
};
template <typename T>
struct alignas(16) SdRec {
    T th;
    float scl;
    int acc;
};

template <typename T, bool WT, int NT = 1>        // NT: coordinate tiles (256 pixels) a pixel buffer holds - 1 except in the long-region objects
struct K2Lay {
    static constexpr int NE = NT * VAMP_TILE_PIX / 2;
    static constexpr int SLOT = NE * (int)sizeof(P2<T>);
    static constexpr int DU = 0, F = SLOT, W = 2 * SLOT, REGION = (WT ? 3 : 2) * SLOT;
    static constexpr int BASE = 0, H = SLOT;
    static __host__ __device__ size_t lsf_bytes(bool lsf) {
        return lsf ? align16(sizeof(T) * (NT * VAMP_TILE_PIX + VAMP_MAX_LSF_TAPS)) : 0;
    }
    static constexpr int SUMSCR = VAMP_K2_SMEMSUM ? 128 : 0;       // 32 floats of reduction scratch, first in the chain slice tail
    static constexpr int RNGSCR = VAMP_K2_SMEMRNG ? 64 * (int)sizeof(P2<T>) : 0;      // the 64 (z, lu) pairs of a Philox block
    static __host__ __device__ size_t bytes(int ncomp_max, bool lsf) {
        return (size_t)H + (size_t)(ncomp_max + 1) * SLOT + lsf_bytes(lsf) + SUMSCR + RNGSCR + sizeof(SdRec<T>) +
               (size_t)ncomp_max * sizeof(CompRec<T>);
    }
};

template <typename T>
struct LsfCtx {
    T* buf;            // warp-private [NT * VAMP_TILE_PIX + taps]
    const T* taps_s;   // block-shared taps
    int taps, P;
    unsigned sum_scratch;   // shared address of the warp's 32-value reduction scratch (VAMP_K2_SMEMSUM)
};

// Sum over the warp through shared memory (VAMP_K2_SMEMSUM, fp32): every lane stores its value, reads all 32 back
// as packed pairs (broadcast reads) and adds them in the same fixed tree, so every lane holds the identical bits -
// as with the shuffle butterfly, whose five dependent shuffle + add rounds this replaces.
template <int OFF>
__device__ __forceinline__ void lds_2x64(unsigned a, P2<float>& x, P2<float>& y) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(x.v), "=l"(y.v) : "r"(a), "n"(OFF) : "memory");
}
__device__ __forceinline__ float warp_sum_smem(float v, unsigned scratch, int lane) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + 4u * (unsigned)lane), "f"(v) : "memory");
    __syncwarp();
    P2<float> q0, q1, q2, q3, q4, q5, q6, q7, q8, q9, qa, qb, qc, qd, qe, qf;
    lds_2x64<0>(scratch, q0, q1);
    lds_2x64<16>(scratch, q2, q3);
    lds_2x64<32>(scratch, q4, q5);
    lds_2x64<48>(scratch, q6, q7);
    lds_2x64<64>(scratch, q8, q9);
    lds_2x64<80>(scratch, qa, qb);
    lds_2x64<96>(scratch, qc, qd);
    lds_2x64<112>(scratch, qe, qf);
    __syncwarp();                                  // the scratch may be rewritten by the next pass
    q0 = p2_add(q0, q1); q2 = p2_add(q2, q3); q4 = p2_add(q4, q5); q6 = p2_add(q6, q7);
    q8 = p2_add(q8, q9); qa = p2_add(qa, qb); qc = p2_add(qc, qd); qe = p2_add(qe, qf);
    q0 = p2_add(q0, q2); q4 = p2_add(q4, q6); q8 = p2_add(q8, qa); qc = p2_add(qc, qe);
    q0 = p2_add(q0, q4); q8 = p2_add(q8, qc);
    q0 = p2_add(q0, q8);
    return p2_lo(q0) + p2_hi(q0);
}
__device__ __forceinline__ double warp_sum_smem(double v, unsigned, int) { return warp_sum(v); }

// Residual sum of squares from a model row-pair buffer in natural order (LSF path only):
// buf[h + p] holds m_p - 1; out = 1 + sum_t lsf[t] (m - 1)[p + h - t], zero outside [0, P).
template <typename T, bool WT, int NT = 1>
__device__ __forceinline__ T lsf_residuals(const LsfCtx<T>& lc, const unsigned char* rb, int npairs, int lane) {
    typedef K2Lay<T, WT, NT> Lay;
    const int h = lc.taps >> 1;
    for (int i = lane; i < 2 * h; i += 32) lc.buf[(i < h) ? i : (64 * npairs + i)] = T(0);
    __syncwarp();
    const T* bq = lc.buf + lane + 2 * h;
    const T* fn = reinterpret_cast<const T*>(rb + Lay::F);
    const T* wn = reinterpret_cast<const T*>(rb + Lay::W);
    T S = T(0);
    for (int j = 0; j < 2 * npairs; ++j) {
        T acc = T(1);
        for (int t = 0; t < lc.taps; ++t) acc = Real<T>::fma_(lc.taps_s[t], bq[-t], acc);
        // the convolution leaks into the padding pixels next to the region edge: mask them
        const int px = j * 32 + lane;
        const int e = 2 * (32 * (j >> 1) + lane) + (j & 1);        // natural pixel -> pair layout
        T rr = (px < lc.P) ? (fn[e] - acc) : T(0);
        if (WT) rr *= wn[e];
        S = Real<T>::fma_(rr, rr, S);
        bq += 32;
    }
    __syncwarp();
    return S;
}

// Shared-memory accessors on 32-bit shared-space addresses (ld/st.shared with an immediate offset).
// The hot passes address every pixel buffer as  base register + compile-time offset : going through C++
// pointers made the compiler re-derive the CTA's shared-window base (S2R SR_CgaCtaId, shifts, LEA, IMAD -
// ~13 instructions) at the top of every pass.
template <typename T> struct SM;
template <> struct SM<float> {
    template <int OFF> static __device__ __forceinline__ P2<float> ld(unsigned a) {
        P2<float> v;
        asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v.v) : "r"(a), "n"(OFF) : "memory");
        return v;
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, P2<float> v) {
        asm volatile("st.shared.b64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "l"(v.v) : "memory");
    }
};
template <> struct SM<double> {
    template <int OFF> static __device__ __forceinline__ P2<double> ld(unsigned a) {
        P2<double> v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.a), "=d"(v.b) : "r"(a), "n"(OFF) : "memory");
        return v;
    }
    template <int OFF> static __device__ __forceinline__ void st(unsigned a, P2<double> v) {
        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v.a), "d"(v.b) : "memory");
    }
};

// Component record (CompRec<T>) through shared addresses: one vector load per field group at the start of a
// component, one read-modify-write at its end.
template <typename T> struct RecIO;
template <> struct RecIO<float> {
    static __device__ __forceinline__ void ld_th(unsigned a, float& t0, float& t1, float& t2, float& t3) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t0), "=f"(t1), "=f"(t2), "=f"(t3) : "r"(a) : "memory");
    }
    static __device__ __forceinline__ void st_th(unsigned a, float t0, float t1, float t2, float t3) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(t0), "f"(t1), "f"(t2), "f"(t3) : "memory");
    }
};
template <> struct RecIO<double> {
    static __device__ __forceinline__ void ld_th(unsigned a, double& t0, double& t1, double& t2, double& t3) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t0), "=d"(t1) : "r"(a) : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(t2), "=d"(t3) : "r"(a) : "memory");
    }
    static __device__ __forceinline__ void st_th(unsigned a, double t0, double t1, double t2, double t3) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(t0), "d"(t1) : "memory");
        asm volatile("st.shared.v2.f64 [%0+16], {%1, %2};" ::"r"(a), "d"(t2), "d"(t3) : "memory");
    }
};
template <int OFF>
__device__ __forceinline__ void sm_ld4(unsigned a, unsigned& x, unsigned& y, unsigned& z, unsigned& w) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sm_st4(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.shared.v4.b32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// One likelihood pass over the region, two row pairs (128 pixels) per trip: half the loop / vote /
// address overhead per pixel of a pair-wise loop and two independent dependency chains per trip.
// FRESH: evaluate the component `cc` (centroid / width steps) and store its shape to the spare slot
// (address ahs); otherwise reuse the cached shape at ahc (amplitude steps, initial state).
// m = ex2(An * H + base); returns sum (w (f - m))^2, identical in every lane.  One path per quad (all
// 128 pixels far -> continued fraction, else Weideman); regions are padded to whole quads with neutral
// pixels (mcmc_chain).  ar / aw: shared addresses of this lane's element 0 in the region slice and in
// the chain slice.  `nfar` counts pixels that took the continued-fraction path (padding included; the
// caller corrects with `last_far`).
// MULT (multiplet regions, vamp_regions_set_multiplet): the component's shape is the sum of its two tied lines,
// H(u; c, g, y) + ar H(u; fr c + dsh, g / fr, y lr / fr) - a second evaluation per fresh quad, everything
// else (cache slots, base, accept) unchanged.  `last_far` then counts the lines of the last quad that took
// the continued fraction (0..2).
template <typename T>
struct Mult2 { T fr, dsh, ar, ginv, yfac; };

// FUSE (cached passes only - the amplitude step, which is the first pass of a component): the move of `base` to this
// component, base += Ap H_prev + Ac H_c (the tau of the previous component comes in, this one's goes out), is done
// inside the pass - one loop, one load of base and H_c instead of two.  Same arithmetic in the same order as the
// separate update loop of mcmc_chain, so the chain is bit-identical.
template <typename T>
struct FuseArgs { T Ap, Ac; unsigned ahp; };

// TAIL (objects built with -DVAMP_INST_TAIL=1; single lines only): `npairs` = ceil(P / 64) row pairs instead of whole
// quads - floor(npairs / 2) quad trips, then ONE pair trip when npairs is odd.  Regions of <= 64 or 129..192 pixels
// stop paying for 64 pixels of padding (BASELINE config C4: P ~ U{32..256}, 14 % of the sampler's lanes).  A chain whose
// pair count is odd ALWAYS takes this path (engine.Sampler sets vamp_mcmc_args.pair_tail when its launch holds one),
// a chain with an even count computes exactly what the quad kernel computes, so results do not depend on the launch.
template <typename T, int KIND, bool WT, bool LSF, bool FRESH, bool MULT = false, bool FUSE = false, bool TAIL = false, int NT = 1>
__device__ __forceinline__ T chain_pass(const CompC<T>& cc, T An, unsigned ahc, unsigned ahs, unsigned ar, unsigned aw,
                                        const unsigned char* rb, int npairs, int lane, const LsfCtx<T>& lc, unsigned& nfar,
                                        unsigned& last_far, const Mult2<T>* ml = nullptr, const FuseArgs<T>* fz = nullptr) {
    typedef K2Lay<T, WT, NT> Lay;
    constexpr int SZ = (int)sizeof(P2<T>);
    constexpr int ROW = 32 * SZ;                       // bytes between the two pairs of a quad
    const P2<T> A2 = p2_bc<T>(An);
    P2<T> S2 = p2_bc<T>(T(0)), S2b = p2_bc<T>(T(0));
    bool far = false, farb = false;
    const unsigned ar_begin = ar;
    const unsigned end = ar + (unsigned)((npairs >> 1) * 2 * ROW);      // end of the whole quads
    unsigned ahp = FUSE ? fz->ahp : 0u;
    const P2<T> Ap2 = p2_bc<T>(FUSE ? fz->Ap : T(0)), Ac2 = p2_bc<T>(FUSE ? fz->Ac : T(0));
    // (without TAIL a region has at least one quad: the loop is bottom-tested, which spares every pass the empty-loop guard)
    if (!TAIL || ar < end)
#pragma unroll K2_TRIPUNROLL
    do {
        P2<T> H0, H1;
        if (FRESH) {
            const P2<T> cb2 = p2_bc<T>(cc.cb), g2 = p2_bc<T>(cc.g);
            const P2<T> X0 = p2_mul(p2_sub(SM<T>::template ld<Lay::DU>(ar), cb2), g2);
            const P2<T> X1 = p2_mul(p2_sub(SM<T>::template ld<Lay::DU + ROW>(ar), cb2), g2);
            if (KIND == VAMP_KIND_GAUSS) {
                const P2<T> Q0 = p2_mul(X0, X0), Q1 = p2_mul(X1, X1);
                H0 = p2_make(Real<T>::exp2_neg(p2_lo(Q0)), Real<T>::exp2_neg(p2_hi(Q0)));
                H1 = p2_make(Real<T>::exp2_neg(p2_lo(Q1)), Real<T>::exp2_neg(p2_hi(Q1)));
            } else {
                const P2<T> y22 = p2_bc<T>(cc.y2);
                const P2<T> S0 = p2_fma(X0, X0, y22), S1 = p2_fma(X1, X1, y22);
                const T smin = fmin(fmin(p2_lo(S0), p2_hi(S0)), fmin(p2_lo(S1), p2_hi(S1)));
                far = __all_sync(VAMP_FULL_MASK, !(smin < Fad<T>::R2()));
                if (far) {
                    nfar += 128u;
                    H0 = fad_far2<T>(X0, cc.y, S0);
                    H1 = fad_far2<T>(X1, cc.y, S1);
                } else {
                    H0 = fad_core2<T>(X0, cc.y, S0);
                    H1 = fad_core2<T>(X1, cc.y, S1);
                }
            }
            if (MULT) {
                // the second line of the component, same two forms
                const T cbb = Real<T>::fma_(ml->fr, cc.cb, ml->dsh), gb = cc.g * ml->ginv, yb = cc.y * ml->yfac;
                const P2<T> cbb2 = p2_bc<T>(cbb), gb2 = p2_bc<T>(gb);
                const P2<T> X0b = p2_mul(p2_sub(SM<T>::template ld<Lay::DU>(ar), cbb2), gb2);
                const P2<T> X1b = p2_mul(p2_sub(SM<T>::template ld<Lay::DU + ROW>(ar), cbb2), gb2);
                P2<T> H0b, H1b;
                if (KIND == VAMP_KIND_GAUSS) {
                    const P2<T> Q0 = p2_mul(X0b, X0b), Q1 = p2_mul(X1b, X1b);
                    H0b = p2_make(Real<T>::exp2_neg(p2_lo(Q0)), Real<T>::exp2_neg(p2_hi(Q0)));
                    H1b = p2_make(Real<T>::exp2_neg(p2_lo(Q1)), Real<T>::exp2_neg(p2_hi(Q1)));
                } else {
                    const P2<T> y22 = p2_bc<T>(yb * yb);
                    const P2<T> S0 = p2_fma(X0b, X0b, y22), S1 = p2_fma(X1b, X1b, y22);
                    const T smin = fmin(fmin(p2_lo(S0), p2_hi(S0)), fmin(p2_lo(S1), p2_hi(S1)));
                    farb = __all_sync(VAMP_FULL_MASK, !(smin < Fad<T>::R2()));
                    if (farb) {
                        nfar += 128u;
                        H0b = fad_far2<T>(X0b, yb, S0);
                        H1b = fad_far2<T>(X1b, yb, S1);
                    } else {
                        H0b = fad_core2<T>(X0b, yb, S0);
                        H1b = fad_core2<T>(X1b, yb, S1);
                    }
                }
                const P2<T> ar2 = p2_bc<T>(ml->ar);
                H0 = p2_fma(ar2, H0b, H0);
                H1 = p2_fma(ar2, H1b, H1);
            }
            SM<T>::template st<0>(ahs, H0);
            SM<T>::template st<ROW>(ahs, H1);
        } else {
            H0 = SM<T>::template ld<0>(ahc);
            H1 = SM<T>::template ld<ROW>(ahc);
        }
        P2<T> B0 = SM<T>::template ld<Lay::BASE>(aw), B1 = SM<T>::template ld<Lay::BASE + ROW>(aw);
        if (FUSE) {
            B0 = p2_fma(Ac2, H0, p2_fma(Ap2, SM<T>::template ld<0>(ahp), B0));
            B1 = p2_fma(Ac2, H1, p2_fma(Ap2, SM<T>::template ld<ROW>(ahp), B1));
            SM<T>::template st<Lay::BASE>(aw, B0);
            SM<T>::template st<Lay::BASE + ROW>(aw, B1);
            ahp += 2 * ROW;
        }
        const P2<T> M0 = p2_exp_scaled<T>(p2_fma(A2, H0, B0));
        const P2<T> M1 = p2_exp_scaled<T>(p2_fma(A2, H1, B1));
        if (LSF) {
            const int e = (int)(ar - ar_begin) / SZ;                      // pair-layout element of this quad
            lc.buf[(lc.taps >> 1) + lane + 2 * e] = p2_lo(M0) - T(1);
            lc.buf[(lc.taps >> 1) + lane + 2 * e + 32] = p2_hi(M0) - T(1);
            lc.buf[(lc.taps >> 1) + lane + 2 * e + 64] = p2_lo(M1) - T(1);
            lc.buf[(lc.taps >> 1) + lane + 2 * e + 96] = p2_hi(M1) - T(1);
        } else {
            P2<T> R0 = p2_sub(SM<T>::template ld<Lay::F>(ar), M0), R1 = p2_sub(SM<T>::template ld<Lay::F + ROW>(ar), M1);
            if (WT) {
                R0 = p2_mul(R0, SM<T>::template ld<Lay::W>(ar));
                R1 = p2_mul(R1, SM<T>::template ld<Lay::W + ROW>(ar));
            }
            S2 = p2_fma(R0, R0, S2);
            S2b = p2_fma(R1, R1, S2b);
        }
        ar += 2 * ROW;
        aw += 2 * ROW;
        ahc += 2 * ROW;
        ahs += 2 * ROW;
    } while (ar < end);
    if (TAIL && (npairs & 1)) {
        // ---- the odd row pair (64 pixels): the quad body for one pair
        P2<T> H0;
        if (FRESH) {
            const P2<T> X0 = p2_mul(p2_sub(SM<T>::template ld<Lay::DU>(ar), p2_bc<T>(cc.cb)), p2_bc<T>(cc.g));
            if (KIND == VAMP_KIND_GAUSS) {
                const P2<T> Q0 = p2_mul(X0, X0);
                H0 = p2_make(Real<T>::exp2_neg(p2_lo(Q0)), Real<T>::exp2_neg(p2_hi(Q0)));
            } else {
                const P2<T> S0 = p2_fma(X0, X0, p2_bc<T>(cc.y2));
                far = __all_sync(VAMP_FULL_MASK, !(fmin(p2_lo(S0), p2_hi(S0)) < Fad<T>::R2()));
                if (far) {
                    nfar += 64u;
                    H0 = fad_far2<T>(X0, cc.y, S0);
                } else {
                    H0 = fad_core2<T>(X0, cc.y, S0);
                }
            }
            SM<T>::template st<0>(ahs, H0);
        } else {
            H0 = SM<T>::template ld<0>(ahc);
        }
        P2<T> B0 = SM<T>::template ld<Lay::BASE>(aw);
        if (FUSE) {
            B0 = p2_fma(Ac2, H0, p2_fma(Ap2, SM<T>::template ld<0>(ahp), B0));
            SM<T>::template st<Lay::BASE>(aw, B0);
        }
        const P2<T> M0 = p2_exp_scaled<T>(p2_fma(A2, H0, B0));
        if (LSF) {
            const int e = (int)(ar - ar_begin) / SZ;
            lc.buf[(lc.taps >> 1) + lane + 2 * e] = p2_lo(M0) - T(1);
            lc.buf[(lc.taps >> 1) + lane + 2 * e + 32] = p2_hi(M0) - T(1);
        } else {
            P2<T> R0 = p2_sub(SM<T>::template ld<Lay::F>(ar), M0);
            if (WT) R0 = p2_mul(R0, SM<T>::template ld<Lay::W>(ar));
            S2 = p2_fma(R0, R0, S2);
        }
    }
    S2 = p2_add(S2, S2b);
    last_far = (far ? 1u : 0u) + ((MULT && farb) ? 1u : 0u);
    T S = p2_lo(S2) + p2_hi(S2);
    if (LSF) S = lsf_residuals<T, WT, NT>(lc, rb, npairs, lane);
#if VAMP_K2_SMEMSUM
    return warp_sum_smem(S, lc.sum_scratch, lane);
#else
    return warp_sum(S);
#endif
}

// ------------------------------------------------------------------------------------------------
// The chain.  PyMC 2's scalar Metropolis-within-Gibbs (oracle/sampler.py; replayed step for step
// by oracle/chain_replay.py): per iteration every component's amplitude, centroid and width(s)
// are stepped in order, then sd.  An amplitude step reuses the cached shape, a centroid / width
// step re-evaluates ONE component over the region (the economy PyMC 2's LazyFunction cache gives
// the reference), `base` carries the other components' tau and moves to the next component with
// two FMAs per pixel.  The step logic runs in the kernel's precision.  Bookkeeping is kept off
// the step path: a component's parameters, scales and counters are one 64-byte record (one
// vector load at the start, one read-modify-write at the end of the component), rejections are
// derived (every parameter is proposed exactly once per iteration) and the post-burn acceptance
// total is settled at tune / burn / exit events.
template <typename T, int KIND, bool WT, bool LSF, bool MULT, bool TAIL = false, int NT = 1>
__device__ void mcmc_chain(const McmcDev& p, int chain, int lane, const unsigned char* rb, unsigned char* wb, int ncomp_max,
                           const T* lsf_s) {
    typedef K2Lay<T, WT, NT> Lay;
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const RegionSetDev& rs = p.rs;
    const int r = p.item_region[chain];
    const int P = rs.npix[r];
    const int n = p.item_ncomp[chain];
    const int ts = p.tstride;
    const int sdi = 4 * p.cmax;
    const int k = NP * n + (WT ? 0 : 1);
    // row pairs swept by every per-pixel loop: with quad trips the region is padded to whole quads (padding
    // pixels - du 1e12, flux 1, weight 0 - contribute exactly nothing)
    const int npairs = TAIL ? ((P + 63) >> 6) : 2 * ((P + 127) >> 7);
    const int npad = 64 * npairs - P;
    const int burn = p.burn;
    const T NK = Real<T>::NEGK();
    const size_t gb = (size_t)chain * ts;

    // ---- carve the slice
    P2<T>* const Hbuf = reinterpret_cast<P2<T>*>(wb + Lay::H) + lane;           // slot s at Hbuf + s * NE
    P2<T>* const pbase = reinterpret_cast<P2<T>*>(wb + Lay::BASE) + lane;
    // 32-bit shared addresses of this lane's element 0 (region slice, chain slice), made opaque so that they
    // stay in registers instead of being re-derived from the shared-window base at every use
    unsigned ar0 = (unsigned)__cvta_generic_to_shared(rb) + (unsigned)(lane * sizeof(P2<T>));
    unsigned aw0 = (unsigned)__cvta_generic_to_shared(wb) + (unsigned)(lane * sizeof(P2<T>));
    asm volatile("" : "+r"(ar0), "+r"(aw0));
    auto slot_addr = [&](int sl) -> unsigned { return aw0 + (unsigned)(Lay::H + sl * Lay::SLOT); };
    unsigned char* sp_ = wb + Lay::H + (size_t)(ncomp_max + 1) * Lay::SLOT;
    LsfCtx<T> lc;
    lc.buf = reinterpret_cast<T*>(sp_);
    lc.taps_s = lsf_s;
    lc.taps = rs.lsf_taps;
    lc.P = P;
    sp_ += Lay::lsf_bytes(LSF);
    lc.sum_scratch = (unsigned)__cvta_generic_to_shared(sp_);
    sp_ += Lay::SUMSCR;
    unsigned rq0 = (unsigned)__cvta_generic_to_shared(sp_);        // VAMP_K2_SMEMRNG: the block's 64 (z, lu) pairs
    sp_ += Lay::RNGSCR;
    SdRec<T>* const sdrec = reinterpret_cast<SdRec<T>*>(sp_);
    CompRec<T>* const recs = reinterpret_cast<CompRec<T>*>(sp_ + sizeof(SdRec<T>));
    unsigned arec0 = (unsigned)__cvta_generic_to_shared(recs);
    asm volatile("" : "+r"(arec0));

    // is parameter i of the flat layout sampled by this chain?
    auto is_free = [&](int i) -> bool { return (i == sdi) ? !WT : ((i < 4 * n) && ((i & 3) < NP)); };

    // ---- state -> shared (the region slice was filled by the CTA)
    for (int e = 0; e < 32 * npairs; e += 32) {
        pbase[e] = p2_bc<T>(T(0));
        Hbuf[n * Lay::NE + e] = p2_bc<T>(T(0));                     // spare slot
    }
    for (int i = lane; i < 4 * n; i += 32) {
        CompRec<T>& rc = recs[i >> 2];
        rc.th[i & 3] = T(p.theta[gb + i]);
        rc.scl[i & 3] = p.asf[gb + i] * p.psd[gb + i];
        rc.acc[i & 3] = p.acc[gb + i];
        if ((i & 3) == 0) rc.slot = i >> 2;
    }
    if (lane == 0) {
        sdrec->th = T(p.theta[gb + sdi]);
        sdrec->scl = p.asf[gb + sdi] * p.psd[gb + sdi];
        sdrec->acc = p.acc[gb + sdi];
    }
    // proposals per parameter since the counters were last reset (the same for every parameter)
    int tot = p.acc[gb] + p.rej[gb];
    if (p.it0 > burn)                        // this launch continues a post-burn counting interval:
        for (int i = lane; i < ts; i += 32) p.acc_total[gb + i] -= p.acc[gb + i];   // settled again at tune / exit
    __syncwarp();

    // centroid relative to the coordinate-tile base; the base is split so the fp32 kernel loses nothing
    const double ubase = rs.ubase[rs.tile_off[r]];
    const T ub_hi = T(ubase), ub_lo = T(ubase - (double)ub_hi);
    auto cbase = [&](T c) -> T { return (c - ub_hi) - ub_lo; };
    // multiplet descriptor of the region, relative to the tile base: cb2 = fr cb + [(fr - 1) ubase + sh]
    Mult2<T> ml;
    ml.fr = T(1);
    ml.dsh = ml.ar = T(0);
    ml.ginv = ml.yfac = T(1);
    if (MULT) {
        MultD md;
        if (region_mult(rs, r, md)) {
            ml.fr = T(md.fr);
            ml.dsh = T((md.fr - 1.0) * ubase + md.sh);
            ml.ar = T(md.ar);
            ml.ginv = T(1.0 / md.fr);
            ml.yfac = T(md.lr / md.fr);
        }
    }

    // ---- build the shape cache; base holds the FULL tau here, one cached pass with A = 0 gives S
    {
        const P2<T>* sd_ = reinterpret_cast<const P2<T>*>(rb + Lay::DU) + lane;
        for (int c = 0; c < n; ++c) {
            CompRec<T>& rc = recs[c];
            const CompC<T> cc = make_comp_fast<T, KIND>(rc.th[0], cbase(rc.th[1]), rc.th[2], rc.th[3]);
            rc.g = (float)cc.g;       // fp32 kernel only (see the centroid step); harmless otherwise
            rc.y = (float)cc.y;
            const P2<T> A2 = p2_bc<T>(NK * cc.A);
            bool far;
            CompC<T> cb_ = cc;                   // second line (MULT)
            cb_.cb = Real<T>::fma_(ml.fr, cc.cb, ml.dsh);
            cb_.g = cc.g * ml.ginv;
            cb_.y = cc.y * ml.yfac;
            cb_.y2 = cb_.y * cb_.y;
            for (int e = 0; e < 32 * npairs; e += 32) {
                P2<T> H = pair_shape<T, KIND>(cc, sd_[e], far);
                if (MULT) H = p2_fma(p2_bc<T>(ml.ar), pair_shape<T, KIND>(cb_, sd_[e], far), H);
                Hbuf[c * Lay::NE + e] = H;
                pbase[e] = p2_fma(A2, H, pbase[e]);
            }
        }
    }
    CompC<T> cz;
    cz.A = cz.cb = cz.g = cz.y = cz.y2 = cz.L2 = T(0);
    unsigned dummy = 0u;
    unsigned dummy_far;
    T Scur = chain_pass<T, KIND, WT, LSF, false, false, false, TAIL, NT>(cz, T(0), slot_addr(n), 0u, ar0, aw0, rb, npairs, lane, lc, dummy, dummy_far);
    T sd = WT ? T(1) : sdrec->th;
    T logsd = WT ? T(0) : StepMath<T>::log_(sd);
    T hiv = WT ? T(0.5) : T(0.5) / (sd * sd);                 // 1 / (2 sd^2)
    T lp = T(0);                                              // sum_c (ln A_c - A_c), kept incrementally
    for (int c = 0; c < n; ++c) {
        const T a = recs[c].th[0];
        lp += StepMath<T>::log_(a) - a;
    }
    const T lognorm = WT ? T(rs.lognorm[r]) : T(0);
    T best_logp = T(p.best_logp[chain]);
    int ntally = p.ntally[chain];
    unsigned long long w_shape = 0ull, w_far = 0ull, w_like = 0ull;

    // ---- random stream: step counter q = 64 qblk + qs
    const unsigned long long uid = p.chain_uid[chain];
    unsigned long long qblk = ((unsigned long long)p.it0 * (unsigned long long)k) >> 6;
    int qs = (int)(((unsigned long long)p.it0 * (unsigned long long)k) & 63ull);
    T zc, luc, zn, lun;                     // current / next half of the 64-step block
    auto refill = [&]() {
#if VAMP_K2_RNG_CALL
        T o4[4];
        rng_refill_call<T>(uid, qblk, (unsigned)lane, p.seed, o4);
        zc = o4[0];
        zn = o4[1];
        luc = o4[2];
        lun = o4[3];
#else
        const uint4 x = philox4x32_10(make_uint4((uint32_t)uid, (uint32_t)(uid >> 32), (uint32_t)qblk,
                                                 ((uint32_t)(qblk >> 32) << 8) | (uint32_t)lane),
                                      make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        RngMath<T>::fill(x, zc, zn, luc, lun);
#endif
    };
#if VAMP_K2_SMEMRNG
    asm volatile("" : "+r"(rq0));
    auto park = [&]() {                     // lane l owns steps l (zc, luc) and 32 + l (zn, lun) of the block
        __syncwarp();
        SM<T>::template st<0>(rq0 + (unsigned)(lane * sizeof(P2<T>)), p2_make(zc, luc));
        SM<T>::template st<0>(rq0 + (unsigned)((32 + lane) * sizeof(P2<T>)), p2_make(zn, lun));
        __syncwarp();
    };
    refill();
    park();
    auto draw = [&](T& z, T& lu) {
        const P2<T> v = SM<T>::template ld<0>(rq0 + (unsigned)(qs * sizeof(P2<T>)));     // broadcast: every lane reads the same pair
        z = p2_lo(v);
        lu = p2_hi(v);
        ++qs;
        if (qs == 64) {
            qs = 0;
            ++qblk;
            refill();
            park();
        }
    };
#else
    refill();
    if (qs >= 32) {
        zc = zn;
        luc = lun;
    }
    auto draw = [&](T& z, T& lu) {
        z = __shfl_sync(VAMP_FULL_MASK, zc, qs);               // source lane = qs mod 32
        lu = __shfl_sync(VAMP_FULL_MASK, luc, qs);
        ++qs;
        if ((qs & 31) == 0) {
            if (qs == 64) {
                qs = 0;
                ++qblk;
                refill();
            } else {
                zc = zn;
                luc = lun;
            }
        }
    };
#endif

    const T c_hi = T(P - 1);
    const T w_hi = T(P - 1) * T(KIND == VAMP_KIND_VOIGT ? 0.5 * FWHM_PER_SIGMA : 0.5);
    // countdowns instead of modulo tests in the iteration loop
    const int tune_interval = p.tune_interval, thin = p.thin;
    int to_tune = (p.it0 % tune_interval == 0) ? 0 : tune_interval - p.it0 % tune_interval;   // iterations until it % ti == 0
    int to_thin = (p.it0 % thin == 0) ? 0 : thin - p.it0 % thin;
    int hs = n;                             // the spare shape slot
    int rebuild = 0;                        // base is rebuilt from the cache when this reaches 0

    for (int it = p.it0; it < p.it1; ++it) {
        // ---- tuning (PyMC 2: every tune_interval iterations; optionally only during burn-in)
        if (to_tune == 0) {
            to_tune = tune_interval;
            if (it > 0 && (p.tune_throughout || it <= burn)) {
                for (int i = lane; i < ts; i += 32) {
                    if (!is_free(i)) continue;
                    int* pa = (i == sdi) ? &sdrec->acc : &recs[i >> 2].acc[i & 3];
                    float* ps = (i == sdi) ? &sdrec->scl : &recs[i >> 2].scl[i & 3];
                    const int a = *pa;
                    float f = p.asf[gb + i];
                    if (tot > 0) {
                        const float nf = f * tune_factor((float)a / (float)tot);
                        if (nf > 0.0f) f = nf;
                    }
                    p.asf[gb + i] = f;
                    *ps = f * p.psd[gb + i];
                    if (it > burn) p.acc_total[gb + i] += a;
                    *pa = 0;
                }
                tot = 0;
                __syncwarp();
            }
        }
        --to_tune;
        if (it == burn) {                    // accepted steps count towards acc_total from here on
            for (int i = lane; i < ts; i += 32)
                if (is_free(i)) p.acc_total[gb + i] -= (i == sdi) ? sdrec->acc : recs[i >> 2].acc[i & 3];
        }
        // work of this iteration: shape evaluations, far-path quads' pixels, passes whose LAST quad (the one holding
        // the padding pixels) went far, in-bounds amplitude passes; settled once per iteration below
        unsigned n_shape = 0u, n_far = 0u, n_lastfar = 0u, n_amp = 0u;
        int hprev = 0;                        // slot of component c - 1
        T Aprev = T(0);                       // its amplitude

#pragma unroll 1
        for (int c = 0; c < n; ++c) {
            constexpr int O_SCL = (int)(4 * sizeof(T)), O_ACC = O_SCL + 16, O_SLOT = O_ACC + 16;     // CompRec<T> layout
            const unsigned arec = arec0 + (unsigned)(c * sizeof(CompRec<T>));
            T pA, pc, pw1, pw2;
            RecIO<T>::ld_th(arec, pA, pc, pw1, pw2);
            unsigned u0, u1, u2, u3, v0, v1, v2, v3;
            sm_ld4<O_SCL>(arec, u0, u1, u2, u3);
            sm_ld4<O_SLOT>(arec, v0, v1, v2, v3);
            const T sA = T(__uint_as_float(u0)), sC = T(__uint_as_float(u1)), sW1 = T(__uint_as_float(u2)), sW2 = T(__uint_as_float(u3));
            int hc = (int)v0;
            // ---- base = tau of the other components (shared-address form, see SM<T>), then the amplitude step on the
            //      cached shape.  VAMP_K2_FUSEBASE: when base only has to MOVE from the previous component to this one
            //      and the amplitude proposal is in bounds, the move is done inside the amplitude pass (FuseArgs).
            unsigned am = 0u;                 // accept bits of this component's steps
#if VAMP_K2_DRAWAHEAD
            T zN, luN;                        // the NEXT step's (normal, log-uniform), drawn one step ahead (same stream order)
#endif
            {
                constexpr int ROW = 32 * (int)sizeof(P2<T>);
                const unsigned endw = aw0 + (unsigned)(npairs * ROW);
                T z, lu;
                draw(z, lu);
                const T prop = pA + z * sA;
#if VAMP_K2_DRAWAHEAD
                draw(zN, luN);                // the centroid step's pair: its shuffles run under the amplitude pass
#endif
                // how base gets to this component: 0 nothing to do (n == 1), 1 rebuild from the cache, 2 move
                int mode = 2;
                T Ap_ = NK * Aprev;
                unsigned ap_ = slot_addr(hprev);
                if (c == 0) {
                    if (rebuild == 0) {
                        mode = 1;
                        rebuild = 8;
                    } else if (n > 1) {
                        // from the last component of the previous iteration back to the first
                        Ap_ = NK * recs[n - 1].th[0];
                        ap_ = slot_addr(recs[n - 1].slot);
                    } else {
                        mode = 0;
                    }
                    --rebuild;
                }
                bool fused = false;
#if VAMP_K2_FUSEBASE && !defined(VAMP_K2_EXPERIMENT_NOAMP_MOVE)
                fused = (mode == 2) && (prop > T(0));
#endif
                if (mode == 1) {
                    // rebuilt from the cache (first iteration, then every 8th) so that rounding never accumulates
                    for (unsigned a = aw0; a < endw; a += ROW) SM<T>::template st<Lay::BASE>(a, p2_bc<T>(T(0)));
                    for (int c2 = 1; c2 < n; ++c2) {
                        const P2<T> A2 = p2_bc<T>(NK * recs[c2].th[0]);
                        unsigned ah = slot_addr(recs[c2].slot);
                        for (unsigned a = aw0; a < endw; a += ROW, ah += ROW)
                            SM<T>::template st<Lay::BASE>(a, p2_fma(A2, SM<T>::template ld<0>(ah), SM<T>::template ld<Lay::BASE>(a)));
                    }
                } else if (mode == 2 && !fused) {
                    // base += tau_prev - tau_c  (two FMAs per pixel instead of n - 1)
                    const P2<T> Ap = p2_bc<T>(Ap_), Ac = p2_bc<T>(-NK * pA);
                    unsigned ap = ap_, an = slot_addr(hc);
                    for (unsigned a = aw0; a < endw; a += ROW, ap += ROW, an += ROW)
                        SM<T>::template st<Lay::BASE>(a, p2_fma(Ac, SM<T>::template ld<0>(an),
                                                               p2_fma(Ap, SM<T>::template ld<0>(ap), SM<T>::template ld<Lay::BASE>(a))));
                }
                // ---- amplitude: cached shape
#ifdef VAMP_K2_EXPERIMENT_NOAMP       // (timing experiment only: what the amplitude step's pass costs)
                if (false) {
#else
                if (prop > T(0)) {
#endif
                    T Snew;
#if VAMP_K2_FUSEBASE
                    if (fused) {
                        FuseArgs<T> fz;
                        fz.Ap = Ap_;
                        fz.Ac = -NK * pA;
                        fz.ahp = ap_;
                        Snew = chain_pass<T, KIND, WT, LSF, false, false, true, TAIL, NT>(cz, NK * prop, slot_addr(hc), 0u, ar0, aw0, rb, npairs,
                                                                                lane, lc, dummy, dummy_far, nullptr, &fz);
                    } else
#endif
                        Snew = chain_pass<T, KIND, WT, LSF, false, false, false, TAIL, NT>(cz, NK * prop, slot_addr(hc), 0u, ar0, aw0, rb, npairs, lane,
                                                                   lc, dummy, dummy_far);
                    const T dprior = StepMath<T>::log_(prop * Real<T>::rcp(pA)) - (prop - pA);
#if !VAMP_K2_NOCOUNT
                    ++n_amp;
#endif
                    if (lu < (Scur - Snew) * hiv + dprior) {          // false for NaN
                        Scur = Snew;
                        pA = prop;
                        lp += dprior;
                        am = 1u;
                    }
                }
            }
            // ---- centroid and width(s): evaluate the proposed shape into the spare slot
#if VAMP_K2_STEPUNROLL
            T cg = T(__uint_as_float(v1)), cy = T(__uint_as_float(v2));     // constants of the current (w1, w2), fp32 kernel
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int i = 1; i < NP; ++i) {
                T z, lu;
#if VAMP_K2_DRAWAHEAD
                z = zN;
                lu = luN;
                if (i + 1 < NP) draw(zN, luN);
#else
                draw(z, lu);
#endif
                const T cur = (i == 1) ? pc : (i == 2) ? pw1 : pw2;
                const T prop = cur + z * ((i == 1) ? sC : (i == 2) ? sW1 : sW2);
                const T hi = (i == 1) ? c_hi : w_hi;
                const bool ok = (prop <= hi) && ((i == NP - 1) ? (prop > T(0)) : (prop >= T(0)));
                if (ok) {
                    const T qc = (i == 1) ? prop : pc, q1 = (i == 2) ? prop : pw1, q2 = (i == 3) ? prop : pw2;
                    CompC<T> cc;
#if VAMP_K2_STEPUNROLL
                    if (i == 1 && sizeof(T) == 4) {       // widths unchanged: reuse g, y (the fp64 kernel recomputes:
                        cc.A = pA;                        // the record holds them in float)
                        cc.cb = cbase(qc);
                        cc.g = cg;
                        cc.y = cy;
                        cc.y2 = cy * cy;
                        cc.L2 = T(0);
                    } else
#endif
                        cc = make_comp_fast<T, KIND>(pA, cbase(qc), q1, q2);
                    unsigned nf = 0u;
                    unsigned last_far;
                    const T Snew = chain_pass<T, KIND, WT, LSF, true, MULT, false, TAIL, NT>(cc, NK * pA, 0u, slot_addr(hs), ar0, aw0, rb, npairs,
                                                                            lane, lc, nf, last_far, &ml);
#if !VAMP_K2_NOCOUNT
                    n_far += nf;
                    n_lastfar += last_far;
                    ++n_shape;
#endif
                    if (lu < (Scur - Snew) * hiv) {
                        Scur = Snew;
                        pc = qc;
                        pw1 = q1;
                        pw2 = q2;
#if VAMP_K2_STEPUNROLL
                        cg = cc.g;
                        cy = cc.y;
#endif
                        const int tmp = hc;         // the proposed shape becomes the component's shape
                        hc = hs;
                        hs = tmp;
                        am |= 1u << i;
                    }
                }
            }
            // ---- component epilogue: every lane writes the same record (same addresses, same values)
            __syncwarp();
            RecIO<T>::st_th(arec, pA, pc, pw1, pw2);
            sm_ld4<O_ACC>(arec, u0, u1, u2, u3);
            sm_st4<O_ACC>(arec, u0 + (am & 1u), u1 + ((am >> 1) & 1u), u2 + ((am >> 2) & 1u), u3 + ((am >> 3) & 1u));
#if VAMP_K2_STEPUNROLL
            sm_st4<O_SLOT>(arec, (unsigned)hc, __float_as_uint((float)cg), __float_as_uint((float)cy), 0u);
#else
            sm_st4<O_SLOT>(arec, (unsigned)hc, v1, v2, 0u);
#endif
            Aprev = pA;
            hprev = hc;
            __syncwarp();
        }
        ++tot;

        if (!WT) {
            // ---- sd: no model evaluation
            T z, lu;
            draw(z, lu);
            const T prop = sd + z * T(sdrec->scl);
            if (prop > T(0) && prop <= T(1)) {
                const T lprop = StepMath<T>::log_(prop);
                const T hv = T(0.5) / (prop * prop);
                if (lu < -Scur * (hv - hiv) - T(P) * (lprop - logsd)) {
                    sd = prop;
                    logsd = lprop;
                    hiv = hv;
                    sdrec->acc += 1;
                    sdrec->th = sd;
                }
            }
            __syncwarp();
        }
        w_shape += (MULT ? 2u : 1u) * n_shape;            // line evaluations over the region
        w_far += n_far - n_lastfar * (unsigned)npad;
        w_like += n_shape + n_amp;

        // ---- end of iteration: log-posterior of the current state
        const T lnL = WT ? (T(-0.5) * Scur + lognorm) : (-Scur * hiv - T(P) * (logsd + T(0.91893853320467274178)));
        const T logp = lnL + lp;
        if (logp > best_logp) {
            best_logp = logp;
            for (int i2 = lane; i2 < ts; i2 += 32)
                if (is_free(i2)) p.best_theta[gb + i2] = (double)((i2 == sdi) ? sdrec->th : recs[i2 >> 2].th[i2 & 3]);
        }
        if (to_thin == 0) {
            to_thin = thin;
            if (it >= burn) {
                auto th_of = [&](int i2) -> double {
                    return is_free(i2) ? (double)((i2 == sdi) ? sdrec->th : recs[i2 >> 2].th[i2 & 3]) : 0.0;
                };
                if (lane == 0) {
                    p.sum_lnL[chain] += (double)lnL;
                    p.sum_lnL2[chain] += (double)lnL * (double)lnL;
                }
                ++ntally;
                const double inv = 1.0 / (double)ntally;
                const double wgt = (double)(ntally - 1) * inv;
                double* mean = p.mean + gb;
                double* m2 = p.m2 + gb;
                double* comom = p.comom + (size_t)chain * 6 * p.cmax;
                for (int i2 = lane; i2 < 6 * n; i2 += 32) {          // co-moments first (need the old means)
                    const int cc2 = i2 / 6, pr = i2 - 6 * cc2;
                    const int a = (pr < 3) ? 0 : (pr < 5 ? 1 : 2);
                    const int b = (pr < 3) ? pr + 1 : (pr < 5 ? pr - 1 : 3);
                    const double da = th_of(VAMP_TH(cc2, a)) - mean[VAMP_TH(cc2, a)];
                    const double db = th_of(VAMP_TH(cc2, b)) - mean[VAMP_TH(cc2, b)];
                    comom[i2] += wgt * da * db;
                }
                __syncwarp();
                for (int i2 = lane; i2 < ts; i2 += 32) {
                    const double x = th_of(i2);
                    const double d = x - mean[i2];
                    const double mn = mean[i2] + d * inv;
                    mean[i2] = mn;
                    m2[i2] += d * (x - mn);
                }
                if (p.trace) {
                    const int slot_t = ntally - 1;
                    if (slot_t < p.trace_cap)
                        for (int i2 = lane; i2 < ts; i2 += 32)
                            p.trace[((size_t)chain * p.trace_cap + slot_t) * ts + i2] = th_of(i2);
                }
                __syncwarp();
            }
        }
        --to_thin;
    }

    // ---- state -> global
    __syncwarp();
    for (int i = lane; i < ts; i += 32) {
        if (!is_free(i)) continue;
        const int a = (i == sdi) ? sdrec->acc : recs[i >> 2].acc[i & 3];
        p.theta[gb + i] = (double)((i == sdi) ? sdrec->th : recs[i >> 2].th[i & 3]);
        p.acc[gb + i] = a;
        p.rej[gb + i] = tot - a;
        if (p.it1 > burn) p.acc_total[gb + i] += a;
    }
    if (lane == 0) {
        const double lnL = WT ? lnL_noise((double)Scur, rs.lognorm[r]) : lnL_sd((double)Scur, (double)sd, P);
        double lpx = 0.0;
        for (int c = 0; c < n; ++c) {
            const double a = (double)recs[c].th[0];
            lpx += log(a) - a;
        }
        p.lnL[chain] = lnL;
        p.logp[chain] = lnL + lpx;
        p.best_logp[chain] = (double)best_logp;
        p.ntally[chain] = ntally;
        if (p.work) {
            p.work[3 * (size_t)chain + 0] += w_shape;    // component evaluations (x P pixels each)
            p.work[3 * (size_t)chain + 1] += w_far;      // pixels evaluated on the continued-fraction path
            p.work[3 * (size_t)chain + 2] += w_like;     // exp(-tau) + residual passes (x P pixels)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Adaptive Metropolis (step_kind = VAMP_STEP_ADAPTIVE; objects built with -DVAMP_INST_AM=1).
// PyMC 2's AdaptiveMetropolis (Haario, Saksman & Tamminen 2001) installed per component - what
// `mcmc.use_step_method(mc.AdaptiveMetropolis, [xexp_c, est_centroid_c, est_L_c, est_G_c])` would do; the reference
// names the method in mcmc_fit (vamp_1.0/vpfits.py:361, 382-389) but never installs it.  Per iteration and
// component ONE joint proposal theta_c + L z (L = Cholesky factor of the block's proposal covariance, z ~ N(0, I)),
// one evaluation of the component's shape, one accept test; sd keeps its scalar Metropolis step with tune().
// Adaptation as in PyMC 2.3: C_0 = diag(|value| / 100); until iteration `delay` only accepted states enter the
// block's history (greedy), afterwards every state; at iteration > delay with it % interval == 0 the covariance is
// updated recursively from the history, scaled by 2.4^2 / d, with epsilon = 1e-5 on the diagonal, and refactored.
// Deviation, stated: the recursion weights the previous covariance by max(k - 1, 0) / (n - 1) (PyMC's (k - 1) is
// negative at the first update, where k = 0), and a covariance that is not positive definite keeps the previous
// factor (PyMC shrinks / repairs it).  Replayed step for step by oracle/chain_replay.py:replay_chain_am.
// The sweep over the region is chain_pass<FRESH> - the arithmetic of the scalar sampler; this path is about the
// step rule, it is not the one bench.py times.
__device__ __forceinline__ int am_tri(int i, int j) { return i * (i + 1) / 2 + j; }

template <typename T, int KIND, bool WT, bool LSF>
__device__ void mcmc_chain_am(const McmcDev& p, int chain, int lane, const unsigned char* rb, unsigned char* wb, int ncomp_max,
                              const T* lsf_s) {
    typedef K2Lay<T, WT> Lay;
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    constexpr int NT = NP * (NP + 1) / 2;
    constexpr int ROW = 32 * (int)sizeof(P2<T>);
    const RegionSetDev& rs = p.rs;
    const int r = p.item_region[chain];
    const int P = rs.npix[r];
    const int n = p.item_ncomp[chain];
    const int ts = p.tstride;
    const int sdi = 4 * p.cmax;
    const int k = NP * n + (WT ? 0 : 1);
    const int npairs = 2 * ((P + 127) >> 7);
    const int npad = 64 * npairs - P;
    const int burn = p.burn;
    const T NK = Real<T>::NEGK();
    const size_t gb = (size_t)chain * ts;
    double* const amg = p.am + (size_t)chain * p.cmax * VAMP_AM_STRIDE;

    P2<T>* const Hbuf = reinterpret_cast<P2<T>*>(wb + Lay::H) + lane;
    P2<T>* const pbase = reinterpret_cast<P2<T>*>(wb + Lay::BASE) + lane;
    unsigned ar0 = (unsigned)__cvta_generic_to_shared(rb) + (unsigned)(lane * sizeof(P2<T>));
    unsigned aw0 = (unsigned)__cvta_generic_to_shared(wb) + (unsigned)(lane * sizeof(P2<T>));
    auto slot_addr = [&](int sl) -> unsigned { return aw0 + (unsigned)(Lay::H + sl * Lay::SLOT); };
    unsigned char* sp_ = wb + Lay::H + (size_t)(ncomp_max + 1) * Lay::SLOT;
    LsfCtx<T> lc;
    lc.buf = reinterpret_cast<T*>(sp_);
    lc.taps_s = lsf_s;
    lc.taps = rs.lsf_taps;
    lc.P = P;
    sp_ += Lay::lsf_bytes(LSF);
    lc.sum_scratch = (unsigned)__cvta_generic_to_shared(sp_);
    sp_ += Lay::SUMSCR;
    unsigned rq0 = (unsigned)__cvta_generic_to_shared(sp_);        // VAMP_K2_SMEMRNG: the block's 64 (z, lu) pairs
    sp_ += Lay::RNGSCR;
    SdRec<T>* const sdrec = reinterpret_cast<SdRec<T>*>(sp_);
    CompRec<T>* const recs = reinterpret_cast<CompRec<T>*>(sp_ + sizeof(SdRec<T>));
    T* const Lsm = reinterpret_cast<T*>(sp_ + sizeof(SdRec<T>) + (size_t)ncomp_max * sizeof(CompRec<T>));   // [ncomp_max][10]
    auto is_free = [&](int i) -> bool { return (i == sdi) ? !WT : ((i < 4 * n) && ((i & 3) < NP)); };

    // ---- state -> shared.  Block counters: acc[0] accepted, acc[1] rejected since the last covariance update
    //      (global: acc / rej of the block's amplitude entry), acc[2] accepted at iterations >= burn of this launch.
    for (int e = 0; e < 32 * npairs; e += 32) {
        pbase[e] = p2_bc<T>(T(0));
        Hbuf[n * Lay::NE + e] = p2_bc<T>(T(0));
    }
    for (int i = lane; i < 4 * n; i += 32) {
        CompRec<T>& rc = recs[i >> 2];
        rc.th[i & 3] = T(p.theta[gb + i]);
        if ((i & 3) == 0) {
            rc.slot = i >> 2;
            rc.acc[0] = p.acc[gb + i];
            rc.acc[1] = p.rej[gb + i];
            rc.acc[2] = 0;
        }
    }
    for (int i = lane; i < 10 * n; i += 32) Lsm[i] = T(amg[(i / 10) * VAMP_AM_STRIDE + 14 + (i % 10)]);
    if (lane == 0) {
        sdrec->th = T(p.theta[gb + sdi]);
        sdrec->scl = p.asf[gb + sdi] * p.psd[gb + sdi];
        sdrec->acc = p.acc[gb + sdi];
    }
    int tot_sd = p.acc[gb + sdi] + p.rej[gb + sdi];
    int sd_post = 0;
    __syncwarp();

    const double ubase = rs.ubase[rs.tile_off[r]];
    const T ub_hi = T(ubase), ub_lo = T(ubase - (double)ub_hi);
    auto cbase = [&](T c) -> T { return (c - ub_hi) - ub_lo; };

    // ---- shape cache and the current sum of squares (as in mcmc_chain)
    {
        const P2<T>* sd_ = reinterpret_cast<const P2<T>*>(rb + Lay::DU) + lane;
        for (int c = 0; c < n; ++c) {
            CompRec<T>& rc = recs[c];
            const CompC<T> cc = make_comp_fast<T, KIND>(rc.th[0], cbase(rc.th[1]), rc.th[2], rc.th[3]);
            const P2<T> A2 = p2_bc<T>(NK * cc.A);
            bool far;
            for (int e = 0; e < 32 * npairs; e += 32) {
                const P2<T> H = pair_shape<T, KIND>(cc, sd_[e], far);
                Hbuf[c * Lay::NE + e] = H;
                pbase[e] = p2_fma(A2, H, pbase[e]);
            }
        }
    }
    CompC<T> cz;
    cz.A = cz.cb = cz.g = cz.y = cz.y2 = cz.L2 = T(0);
    unsigned dummy = 0u, dummy_far;
    T Scur = chain_pass<T, KIND, WT, LSF, false>(cz, T(0), slot_addr(n), 0u, ar0, aw0, rb, npairs, lane, lc, dummy, dummy_far);
    T sd = WT ? T(1) : sdrec->th;
    T logsd = WT ? T(0) : StepMath<T>::log_(sd);
    T hiv = WT ? T(0.5) : T(0.5) / (sd * sd);
    T lp = T(0);
    for (int c = 0; c < n; ++c) {
        const T a = recs[c].th[0];
        lp += StepMath<T>::log_(a) - a;
    }
    const T lognorm = WT ? T(rs.lognorm[r]) : T(0);
    T best_logp = T(p.best_logp[chain]);
    int ntally = p.ntally[chain];
    unsigned long long w_shape = 0ull, w_far = 0ull, w_like = 0ull;

    // ---- random stream: the scalar sampler's layout (k draws per iteration); a block uses the normals of its NP
    //      draws and the log-uniform of the first
    const unsigned long long uid = p.chain_uid[chain];
    unsigned long long qblk = ((unsigned long long)p.it0 * (unsigned long long)k) >> 6;
    int qs = (int)(((unsigned long long)p.it0 * (unsigned long long)k) & 63ull);
    T zc, luc, zn, lun;
    auto refill = [&]() {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)uid, (uint32_t)(uid >> 32), (uint32_t)qblk,
                                                 ((uint32_t)(qblk >> 32) << 8) | (uint32_t)lane),
                                      make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        RngMath<T>::fill(x, zc, zn, luc, lun);
    };
    refill();
    if (qs >= 32) {
        zc = zn;
        luc = lun;
    }
    auto draw = [&](T& z, T& lu) {
        z = __shfl_sync(VAMP_FULL_MASK, zc, qs);
        lu = __shfl_sync(VAMP_FULL_MASK, luc, qs);
        ++qs;
        if ((qs & 31) == 0) {
            if (qs == 64) {
                qs = 0;
                ++qblk;
                refill();
            } else {
                zc = zn;
                luc = lun;
            }
        }
    };

    const T c_hi = T(P - 1);
    const T w_hi = T(P - 1) * T(KIND == VAMP_KIND_VOIGT ? 0.5 * FWHM_PER_SIGMA : 0.5);
    const int tune_interval = p.tune_interval, thin = p.thin;
    int hs = n;
    const unsigned endw = aw0 + (unsigned)(npairs * ROW);

    for (int it = p.it0; it < p.it1; ++it) {
        // ---- sd: PyMC 2 Metropolis.tune every tune_interval iterations
        if (!WT && it > 0 && it % tune_interval == 0 && (p.tune_throughout || it <= burn)) {
            if (lane == 0) {
                float f = p.asf[gb + sdi];
                if (tot_sd > 0) {
                    const float nf = f * tune_factor((float)sdrec->acc / (float)tot_sd);
                    if (nf > 0.0f) f = nf;
                }
                p.asf[gb + sdi] = f;
                sdrec->scl = f * p.psd[gb + sdi];
                sdrec->acc = 0;
            }
            tot_sd = 0;
            __syncwarp();
        }
        unsigned n_shape = 0u, n_far = 0u, n_like = 0u;
        int hprev = 0;
        T Aprev = T(0);
        const bool post_delay = it >= p.am_delay;              // PyMC: greedy is switched off AT iteration == delay

        for (int c = 0; c < n; ++c) {
            CompRec<T>& rc = recs[c];
            T th[4] = {rc.th[0], rc.th[1], rc.th[2], rc.th[3]};
            int hc = rc.slot;
            // ---- base = tau of the other components
            if (c == 0) {
                for (unsigned a = aw0; a < endw; a += ROW) SM<T>::template st<Lay::BASE>(a, p2_bc<T>(T(0)));
                for (int c2 = 1; c2 < n; ++c2) {
                    const P2<T> A2 = p2_bc<T>(NK * recs[c2].th[0]);
                    unsigned ah = slot_addr(recs[c2].slot);
                    for (unsigned a = aw0; a < endw; a += ROW, ah += ROW)
                        SM<T>::template st<Lay::BASE>(a, p2_fma(A2, SM<T>::template ld<0>(ah), SM<T>::template ld<Lay::BASE>(a)));
                }
            } else {
                const P2<T> Ap = p2_bc<T>(NK * Aprev), Ac = p2_bc<T>(-NK * th[0]);
                unsigned ap = slot_addr(hprev), an = slot_addr(hc);
                for (unsigned a = aw0; a < endw; a += ROW, ap += ROW, an += ROW)
                    SM<T>::template st<Lay::BASE>(a, p2_fma(Ac, SM<T>::template ld<0>(an),
                                                           p2_fma(Ap, SM<T>::template ld<0>(ap), SM<T>::template ld<Lay::BASE>(a))));
            }
            // ---- joint proposal theta + L z
            T z[4] = {T(0), T(0), T(0), T(0)}, lu = T(0);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                T lui;
                draw(z[i], lui);
                if (i == 0) lu = lui;
            }
            const T* Lc = Lsm + 10 * c;
            T prop[4] = {th[0], th[1], th[2], th[3]};
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) prop[i] = Real<T>::fma_(Lc[am_tri(i, j)], z[j], prop[i]);
            bool ok = (prop[0] > T(0)) && (prop[1] >= T(0)) && (prop[1] <= c_hi);
            if (KIND == VAMP_KIND_VOIGT)
                ok = ok && (prop[2] >= T(0)) && (prop[2] <= w_hi) && (prop[3] > T(0)) && (prop[3] <= w_hi);
            else
                ok = ok && (prop[2] > T(0)) && (prop[2] <= w_hi);
            bool accept = false;
            if (ok) {
                const CompC<T> cc = make_comp_fast<T, KIND>(prop[0], cbase(prop[1]), prop[2], prop[3]);
                unsigned nf = 0u, last_far;
                const T Snew = chain_pass<T, KIND, WT, LSF, true>(cc, NK * prop[0], 0u, slot_addr(hs), ar0, aw0, rb, npairs, lane, lc,
                                                                  nf, last_far);
                n_far += nf - last_far * (unsigned)npad;
                ++n_shape;
                ++n_like;
                const T dprior = StepMath<T>::log_(prop[0] * Real<T>::rcp(th[0])) - (prop[0] - th[0]);
                if (lu < (Scur - Snew) * hiv + dprior) {
                    accept = true;
                    Scur = Snew;
                    lp += dprior;
#pragma unroll
                    for (int i = 0; i < NP; ++i) th[i] = prop[i];
                    const int tmp = hc;
                    hc = hs;
                    hs = tmp;
                }
            }
            __syncwarp();
            if (lane == 0) {
                rc.th[0] = th[0];
                rc.th[1] = th[1];
                rc.th[2] = th[2];
                rc.th[3] = th[3];
                rc.slot = hc;
                rc.acc[accept ? 0 : 1] += 1;
                if (accept && it >= burn) rc.acc[2] += 1;
            }
            // ---- the block's history (greedy until `delay`): sum x, sum x x^T, count
            if (accept || post_delay) {
                double* a = amg + c * VAMP_AM_STRIDE;
                if (lane < NP) {
                    a[24 + lane] += (double)th[lane];
                } else if (lane < NP + NT) {
                    int e = lane - NP, i = 0;
                    while (am_tri(i + 1, 0) <= e) ++i;
                    const int j = e - am_tri(i, 0);
                    a[28 + e] += (double)th[i] * (double)th[j];
                } else if (lane == 31) {
                    a[39] += 1.0;
                }
            }
            Aprev = th[0];
            hprev = hc;
            __syncwarp();
        }

        if (!WT) {
            // ---- sd: scalar Metropolis, no model evaluation
            T z, lu;
            draw(z, lu);
            const T prop = sd + z * T(sdrec->scl);
            bool acc_sd = false;
            if (prop > T(0) && prop <= T(1)) {
                const T lprop = StepMath<T>::log_(prop);
                const T hv = T(0.5) / (prop * prop);
                if (lu < -Scur * (hv - hiv) - T(P) * (lprop - logsd)) {
                    sd = prop;
                    logsd = lprop;
                    hiv = hv;
                    acc_sd = true;
                }
            }
            __syncwarp();
            if (acc_sd && lane == 0) {
                sdrec->acc += 1;
                sdrec->th = sd;
            }
            if (acc_sd && it >= burn) ++sd_post;
            ++tot_sd;
            __syncwarp();
        }
        w_shape += n_shape;
        w_far += n_far;
        w_like += n_like;

        // ---- covariance update (PyMC 2: _current_iter > delay and _current_iter % interval == 0)
        if (it > p.am_delay && it % p.am_interval == 0) {
            __syncwarp();
            for (int c = lane; c < n; c += 32) {
                double* a = amg + c * VAMP_AM_STRIDE;
                const double kk = a[38], m = a[39], nn = kk + m;
                if (m > 0.0 && nn > 1.0) {
                    double nm[4];
                    for (int i = 0; i < NP; ++i) nm[i] = (kk * a[i] + a[24 + i]) / nn;
                    const double scaling = 5.76 / (double)NP, eps = 1.0e-5, wold = (kk > 1.0 ? kk - 1.0 : 0.0) / (nn - 1.0);
                    for (int i = 0; i < NP; ++i)
                        for (int j = 0; j <= i; ++j) {
                            const int e = am_tri(i, j);
                            const double t = kk * a[i] * a[j] + a[28 + e] - nn * nm[i] * nm[j] + (i == j ? eps : 0.0);
                            a[4 + e] = wold * a[4 + e] + scaling / (nn - 1.0) * t;
                        }
                    for (int i = 0; i < NP; ++i) {
                        a[i] = nm[i];
                        a[24 + i] = 0.0;
                    }
                    for (int e = 0; e < NT; ++e) a[28 + e] = 0.0;
                    a[38] = nn;
                    a[39] = 0.0;
                    double Lf[10];
                    bool pd = true;
                    for (int i = 0; i < NP && pd; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double s_ = a[4 + am_tri(i, j)];
                            for (int q = 0; q < j; ++q) s_ -= Lf[am_tri(i, q)] * Lf[am_tri(j, q)];
                            if (i == j) {
                                if (!(s_ > 0.0)) {
                                    pd = false;
                                    break;
                                }
                                Lf[am_tri(i, i)] = sqrt(s_);
                            } else {
                                Lf[am_tri(i, j)] = s_ / Lf[am_tri(j, j)];
                            }
                        }
                    if (pd)
                        for (int e = 0; e < NT; ++e) {
                            a[14 + e] = Lf[e];
                            Lsm[10 * c + e] = T(Lf[e]);
                        }
                }
                recs[c].acc[0] = 0;
                recs[c].acc[1] = 0;
            }
            __syncwarp();
        }

        // ---- end of iteration: log-posterior, best state, tallies (as in mcmc_chain)
        const T lnL = WT ? (T(-0.5) * Scur + lognorm) : (-Scur * hiv - T(P) * (logsd + T(0.91893853320467274178)));
        const T logp = lnL + lp;
        if (logp > best_logp) {
            best_logp = logp;
            for (int i2 = lane; i2 < ts; i2 += 32)
                if (is_free(i2)) p.best_theta[gb + i2] = (double)((i2 == sdi) ? sdrec->th : recs[i2 >> 2].th[i2 & 3]);
        }
        if (it >= burn && it % thin == 0) {
            auto th_of = [&](int i2) -> double {
                return is_free(i2) ? (double)((i2 == sdi) ? sdrec->th : recs[i2 >> 2].th[i2 & 3]) : 0.0;
            };
            if (lane == 0) {
                p.sum_lnL[chain] += (double)lnL;
                p.sum_lnL2[chain] += (double)lnL * (double)lnL;
            }
            ++ntally;
            const double inv = 1.0 / (double)ntally;
            const double wgt = (double)(ntally - 1) * inv;
            double* mean = p.mean + gb;
            double* m2 = p.m2 + gb;
            double* comom = p.comom + (size_t)chain * 6 * p.cmax;
            for (int i2 = lane; i2 < 6 * n; i2 += 32) {
                const int cc2 = i2 / 6, pr = i2 - 6 * cc2;
                const int a = (pr < 3) ? 0 : (pr < 5 ? 1 : 2);
                const int b = (pr < 3) ? pr + 1 : (pr < 5 ? pr - 1 : 3);
                const double da = th_of(VAMP_TH(cc2, a)) - mean[VAMP_TH(cc2, a)];
                const double db = th_of(VAMP_TH(cc2, b)) - mean[VAMP_TH(cc2, b)];
                comom[i2] += wgt * da * db;
            }
            __syncwarp();
            for (int i2 = lane; i2 < ts; i2 += 32) {
                const double x = th_of(i2);
                const double d = x - mean[i2];
                const double mn = mean[i2] + d * inv;
                mean[i2] = mn;
                m2[i2] += d * (x - mn);
            }
            if (p.trace) {
                const int slot_t = ntally - 1;
                if (slot_t < p.trace_cap)
                    for (int i2 = lane; i2 < ts; i2 += 32)
                        p.trace[((size_t)chain * p.trace_cap + slot_t) * ts + i2] = th_of(i2);
            }
            __syncwarp();
        }
    }

    // ---- state -> global
    __syncwarp();
    for (int i = lane; i < ts; i += 32) {
        if (!is_free(i)) continue;
        p.theta[gb + i] = (double)((i == sdi) ? sdrec->th : recs[i >> 2].th[i & 3]);
        if (i == sdi) {
            p.acc[gb + i] = sdrec->acc;
            p.rej[gb + i] = tot_sd - sdrec->acc;
            p.acc_total[gb + i] += sd_post;
        } else if ((i & 3) == 0) {
            p.acc[gb + i] = recs[i >> 2].acc[0];
            p.rej[gb + i] = recs[i >> 2].acc[1];
            p.acc_total[gb + i] += recs[i >> 2].acc[2];
        }
    }
    if (lane == 0) {
        const double lnL = WT ? lnL_noise((double)Scur, rs.lognorm[r]) : lnL_sd((double)Scur, (double)sd, P);
        double lpx = 0.0;
        for (int c = 0; c < n; ++c) {
            const double a = (double)recs[c].th[0];
            lpx += log(a) - a;
        }
        p.lnL[chain] = lnL;
        p.logp[chain] = lnL + lpx;
        p.best_logp[chain] = (double)best_logp;
        p.ntally[chain] = ntally;
        if (p.work) {
            p.work[3 * (size_t)chain + 0] += w_shape;
            p.work[3 * (size_t)chain + 1] += w_far;
            p.work[3 * (size_t)chain + 2] += w_like;
        }
    }
}

// One warp per chain, K2_WARPS chains per CTA.  The host sorts chains by (pixel class, component
// count, model), so the warps of a CTA almost always fit the same region: the region's pixels
// (du, flux, 1/noise) are staged ONCE per CTA and shared, which is what lets a sixth CTA fit on an
// SM.  A CTA whose warps belong to different regions (group boundaries) processes its distinct
// regions one after the other - rare, and always correct.
#ifdef VAMP_K2_MAXNREG              // experiments with other CTA sizes: a plain register cap instead of launch bounds
#define VAMP_K2_BOUNDS __maxnreg__(VAMP_K2_MAXNREG)
#else
#define VAMP_K2_BOUNDS __launch_bounds__(KW * 32, (NT > 1) ? 1 : (sizeof(T) == 4) ? VAMP_K2_MINBLOCKS : VAMP_K2_MINBLOCKS64)
#endif
// NT > 1 (long-region objects, -DVAMP_INST_LONG=1): regions of up to NT coordinate tiles.  Every pixel buffer is NT
// tiles long, KW = 1 chain per CTA, and the region slice is staged with ALL its tiles relative to the base of the
// first one (the sum is formed in float64 and rounded once: in float32 a coordinate 2048 pixels from the base keeps
// 1.2e-4 pixel, far below any proposal scale), so the chain itself is the single-tile code.
template <typename T, int KIND, bool WT, bool LSF, bool MULT, bool AM = false, bool TAIL = false, int NT = 1, int KW = K2_WARPS>
__global__ void VAMP_K2_BOUNDS
k2_mcmc_kernel(McmcDev p, int ncomp_max) {
    typedef K2Lay<T, WT, NT> Lay;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_region[KW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const size_t head = LSF ? align16(sizeof(T) * VAMP_MAX_LSF_TAPS) : 0;
    if (LSF) {
        for (int i = threadIdx.x; i < p.rs.lsf_taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(p.rs)[i];
    }
    // launch order: the host sorts chains by ascending cost; the longest chains go first so the
    // tail of the grid is made of short ones
    const int chain = p.B - 1 - (blockIdx.x * KW + warp);
    const int my_region = (chain >= 0) ? p.item_region[chain] : -1;
    if (lane == 0) s_region[warp] = my_region;
    __syncthreads();
    unsigned char* rb = smem_raw + head;
    unsigned char* wb = rb + Lay::REGION + (size_t)warp * (Lay::bytes(ncomp_max, LSF) + (AM ? align16((size_t)ncomp_max * 10 * sizeof(T)) : 0));
    for (int round = 0; round < KW; ++round) {
        const int rr = s_region[round];
        bool skip = rr < 0;
        for (int j = 0; j < round; ++j) skip = skip || (s_region[j] == rr);
        if (skip) continue;                              // CTA-uniform
        if (round > 0) __syncthreads();                  // the previous region's chains are done
        if (p.rs.npix[rr] > NT * VAMP_TILE_PIX) __trap();  // a caller's vamp_mcmc_args.max_npix below its longest region: fail loudly
        {
            const RegionSetDev& rs = p.rs;
            const int off = rs.pix_off[rr] >> 1, ne = 32 * ((rs.npix[rr] + 63) >> 6);
            const int nq = TAIL ? ne : 64 * ((rs.npix[rr] + 127) >> 7);               // whole quads: pad the rest
            P2<T>* sd_ = reinterpret_cast<P2<T>*>(rb + Lay::DU);
            P2<T>* sf_ = reinterpret_cast<P2<T>*>(rb + Lay::F);
            P2<T>* sw_ = reinterpret_cast<P2<T>*>(rb + Lay::W);
            for (int e = threadIdx.x; e < nq; e += blockDim.x) {
                const bool in = e < ne;
                sd_[e] = in ? p2_from(__ldg(RS<T>::pdu(rs) + off + e)) : p2_bc<T>(T(1.0e12));
                if (NT > 1 && in) {
                    // element e = 32 jj + lane holds pixels 64 jj + lane (+ 32): tile jj >> 2; its coordinates move to tile 0's base
                    const double sh = rs.ubase[rs.tile_off[rr] + (e >> 7)] - rs.ubase[rs.tile_off[rr]];
                    sd_[e] = p2_make(T((double)p2_lo(sd_[e]) + sh), T((double)p2_hi(sd_[e]) + sh));
                }
                sf_[e] = in ? p2_from(__ldg(RS<T>::pf(rs) + off + e)) : p2_bc<T>(T(1));
                if (WT) sw_[e] = in ? p2_from(__ldg(RS<T>::pw(rs) + off + e)) : p2_bc<T>(T(0));
            }
        }
        __syncthreads();
        if (my_region == rr) {
            if (AM) mcmc_chain_am<T, KIND, WT, LSF>(p, chain, lane, rb, wb, ncomp_max, lsf_s);
            else mcmc_chain<T, KIND, WT, LSF, MULT, TAIL, NT>(p, chain, lane, rb, wb, ncomp_max, lsf_s);
        }
    }
}

// ---- initial states -------------------------------------------------------------------------
// PyMC 2 semantics (oracle/sampler.py:Model.initial): xexp value = 0.5 (vpfits.py:240,284); a
// mc.Uniform without value= starts at a draw from its prior; sd ~ U(0,1) likewise.
template <int KIND>
__global__ void init_prior_kernel(McmcDev p) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= p.B) return;
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const int P = p.rs.npix[p.item_region[chain]];
    const int n = p.item_ncomp[chain];
    const double c_hi = (double)(P - 1);
    const double w_hi = 0.5 * (double)(P - 1) * (KIND == VAMP_KIND_VOIGT ? FWHM_PER_SIGMA : 1.0);
    const unsigned long long uid = p.chain_uid[chain];
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32) ^ 0x494E4954u);   // "INIT" stream
    double* th = p.theta + (size_t)chain * p.tstride;
    for (int i = 0; i < p.tstride; ++i) th[i] = 0.0;
    for (int c = 0; c < n; ++c) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)uid, (uint32_t)(uid >> 32), (uint32_t)c, 0u), key);
        th[VAMP_TH(c, 0)] = 0.5;
        th[VAMP_TH(c, 1)] = u01(x.x) * c_hi;
        th[VAMP_TH(c, 2)] = u01(x.y) * w_hi;
        if (NP == 4) th[VAMP_TH(c, 3)] = u01(x.z) * w_hi;
    }
    if (p.sd_free) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)uid, (uint32_t)(uid >> 32), 0xFFFFFFFFu, 0u), key);
        th[4 * p.cmax] = u01(x.x);
    }
}

// Reset sampler state around a given theta: proposal_sd = |value| (1 if 0), asf = 1, counters 0.
__global__ void init_state_kernel(McmcDev p) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= p.B) return;
    const size_t gb = (size_t)chain * p.tstride;
    for (int i = 0; i < p.tstride; ++i) {
        const double v = p.theta[gb + i];
        p.psd[gb + i] = (v != 0.0) ? (float)fabs(v) : 1.0f;
        p.asf[gb + i] = 1.0f;
        p.acc[gb + i] = 0;
        p.rej[gb + i] = 0;
        p.acc_total[gb + i] = 0;
        p.best_theta[gb + i] = v;
        p.mean[gb + i] = 0.0;
        p.m2[gb + i] = 0.0;
    }
    for (int i = 0; i < 6 * p.cmax; ++i) p.comom[(size_t)chain * 6 * p.cmax + i] = 0.0;
    if (p.am) {
        // PyMC 2 AdaptiveMetropolis.cov_from_value(100): C_0 = diag(|value| / 100) (1 for a zero value); empty history
        for (int c = 0; c < p.cmax; ++c) {
            double* a = p.am + ((size_t)chain * p.cmax + c) * VAMP_AM_STRIDE;
            for (int i = 0; i < VAMP_AM_STRIDE; ++i) a[i] = 0.0;
            for (int i = 0; i < 4; ++i) {
                const double v = p.theta[gb + 4 * c + i];
                const double var = ((v != 0.0) ? fabs(v) : 1.0) / 100.0;
                a[4 + i * (i + 1) / 2 + i] = var;
                a[14 + i * (i + 1) / 2 + i] = sqrt(var);
            }
        }
    }
    p.best_logp[chain] = -INFINITY;
    p.logp[chain] = -INFINITY;
    p.lnL[chain] = -INFINITY;
    p.ntally[chain] = 0;
    p.sum_lnL[chain] = 0.0;
    p.sum_lnL2[chain] = 0.0;
}

#ifndef VAMP_INST_AM
#define VAMP_INST_AM 0            // 1: this object holds the adaptive-Metropolis instantiations (vamp_launch_k2a_*)
#endif
#ifndef VAMP_INST_TAIL
#define VAMP_INST_TAIL 0          // 1: this object holds the pair-tail instantiations (vamp_launch_k2t_*)
#endif
#ifndef VAMP_INST_LONG
#define VAMP_INST_LONG 0          // 1: this object holds the long-region instantiations (vamp_launch_k2l_*): 257 .. 2048 pixels
#endif

template <typename T, int KIND, bool WT, bool LSF, int NT = 1, int KW = K2_WARPS>
cudaError_t launch_k2_t(const McmcDev& p, int max_ncomp, cudaStream_t st) {
    constexpr bool AM = VAMP_INST_AM != 0;
    const size_t smem = (LSF ? align16(sizeof(T) * VAMP_MAX_LSF_TAPS) : 0) + K2Lay<T, WT, NT>::REGION +
                        KW * (K2Lay<T, WT, NT>::bytes(max_ncomp, LSF) + (AM ? align16((size_t)max_ncomp * 10 * sizeof(T)) : 0));
    auto kern = k2_mcmc_kernel<T, KIND, WT, LSF, VAMP_INST_MULT != 0, AM, VAMP_INST_TAIL != 0, NT, KW>;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (p.B + KW - 1) / KW;
    kern<<<grid, KW * 32, smem, st>>>(p, max_ncomp);
    return cudaGetLastError();
}

}  // namespace

#if VAMP_INST_F64 == 0 && VAMP_INST_MULT == 0 && VAMP_INST_AM == 0 && VAMP_INST_TAIL == 0 && VAMP_INST_LONG == 0
// precision-independent helpers live in the f32 objects only
cudaError_t VAMP_CAT3(vamp_launch_init_prior, any, VAMP_KN)(const McmcDev& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    init_prior_kernel<InstKind><<<(p.B + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}
#if VAMP_INST_VOIGT == 0
cudaError_t vamp_launch_init_state(const McmcDev& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    init_state_kernel<<<(p.B + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}
#endif
#endif

// -DVAMP_INST_MULT=1 objects hold the multiplet instantiations (vamp_launch_k2m_*), chosen by the C ABI when the
// region set carries a multiplet table
#if VAMP_INST_LONG
// Regions of 257 .. 2048 pixels (the reference fits whole arrays of any length, e.g. 2000 pixels in vpfits_intro.ipynb):
// single lines, scalar Metropolis.  Two buffer lengths, <= 512 and <= 2048 pixels; shared memory is
// (max_ncomp + 4) x NT KB per chain (float32), so the launch fails with cudaErrorInvalidConfiguration beyond
// max_ncomp = 52 (1024 pixels, float32) / 24 (2048 pixels) - vamp_mcmc_run reports it.
template <int NT>
static cudaError_t launch_k2_long(const McmcDev& p, int max_ncomp, cudaStream_t st) {
    const bool lsf = p.rs.lsf_taps > 1;
    if (p.sd_free)
        return lsf ? launch_k2_t<InstT, InstKind, false, true, NT, 1>(p, max_ncomp, st)
                   : launch_k2_t<InstT, InstKind, false, false, NT, 1>(p, max_ncomp, st);
    return lsf ? launch_k2_t<InstT, InstKind, true, true, NT, 1>(p, max_ncomp, st)
               : launch_k2_t<InstT, InstKind, true, false, NT, 1>(p, max_ncomp, st);
}
cudaError_t VAMP_INST_NAME(vamp_launch_k2l)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    if (p.B <= 0 || p.it1 <= p.it0) return cudaSuccess;
    if (max_npix > VAMP_K2_LONG_PIX || p.rs.mult) return cudaErrorInvalidValue;
    if (max_npix <= 2 * VAMP_TILE_PIX) return launch_k2_long<2>(p, max_ncomp, st);
    if (max_npix <= 4 * VAMP_TILE_PIX) return launch_k2_long<4>(p, max_ncomp, st);
    return launch_k2_long<VAMP_K2_LONG_PIX / VAMP_TILE_PIX>(p, max_ncomp, st);
}
#else
#if VAMP_INST_AM
cudaError_t VAMP_INST_NAME(vamp_launch_k2a)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    if (!p.am) return cudaErrorInvalidValue;
#elif VAMP_INST_TAIL
cudaError_t VAMP_INST_NAME(vamp_launch_k2t)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
#elif VAMP_INST_MULT
cudaError_t VAMP_INST_NAME(vamp_launch_k2m)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
#else
cudaError_t VAMP_INST_NAME(vamp_launch_k2)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
#endif
    if (p.B <= 0 || p.it1 <= p.it0) return cudaSuccess;
    if (max_npix > VAMP_TILE_PIX) return cudaErrorInvalidValue;
    const bool lsf = p.rs.lsf_taps > 1;
    if (p.sd_free)
        return lsf ? launch_k2_t<InstT, InstKind, false, true>(p, max_ncomp, st)
                   : launch_k2_t<InstT, InstKind, false, false>(p, max_ncomp, st);
    return lsf ? launch_k2_t<InstT, InstKind, true, true>(p, max_ncomp, st)
               : launch_k2_t<InstT, InstKind, true, false>(p, max_ncomp, st);
}
#endif   // VAMP_INST_LONG
