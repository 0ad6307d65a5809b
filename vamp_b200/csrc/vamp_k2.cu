// vamp_k2.cu - K2: batched Metropolis sampler, one warp per chain, whole chains on device.
//
// Replaces the reference's hot loop  mcmc.sample(iter, burn, thin)  (vamp_1.0/vpfits.py:393,
// 423-425): PyMC 2's auto-assigned scalar Metropolis steppers with adaptive-scale tuning
// (restated in oracle/sampler.py).  Thousands of independent chains - every candidate
// component count of the BIC sweep and every restart - advance inside ONE launch; there is no
// host round trip, no kernel-per-iteration, and no HBM traffic beyond the initial spectrum
// load and the optional tallies.
//
// Per chain (= warp):
//   * pixels p = j*32 + lane live in registers (flux, coordinate, running tau);
//   * the unit-amplitude shape H_c of every component is cached in warp-private shared
//     memory, so a single-parameter step re-evaluates ONE component (P Voigt evaluations) -
//     the same economy PyMC 2's LazyFunction cache gives the reference - and an amplitude or
//     noise-parameter step re-evaluates none;
//   * sum (f-m)^2 is reduced with xor shuffles: every lane holds the identical value and takes
//     the identical accept decision (no divergence, no broadcast);
//   * Philox4x32-10 keyed by (seed, global chain id): lane l draws the Box-Muller pair and the
//     two log-uniforms of steps 2l, 2l+1 of a 64-step block, handed out with shuffles;
//   * tallies (i >= burn and i % thin == 0) feed Welford means / co-moments (the reference's
//     mcmc.stats() and chain_covariance, vpfits.py:432-456) and an optional trace.
#include "vamp_device.cuh"
#include "vamp_kernels.h"

namespace {

constexpr int K2_WARPS = 4;
#ifndef VAMP_K2_MINBLOCKS
#define VAMP_K2_MINBLOCKS 5
#endif
constexpr double FWHM_PER_SIGMA = 2.35482004503094938202;

template <typename T>
struct WarpRng {
    T z0, z1, lu0, lu1;      // kept in the kernel's precision: half the registers in the fp32 kernel
    uint32_t uid_lo, uid_hi, k0, k1;
    __device__ __forceinline__ void refill(unsigned long long blk, int lane) {
        const uint4 x = philox4x32_10(make_uint4(uid_lo, uid_hi, (uint32_t)blk, ((uint32_t)(blk >> 32) << 8) | (uint32_t)lane),
                                      make_uint2(k0, k1));
        const double rad = sqrt(-2.0 * log(u01(x.x)));
        double sn, cs;
        sincospi(2.0 * u01(x.y), &sn, &cs);
        z0 = T(rad * cs);
        z1 = T(rad * sn);
        lu0 = T(log(u01(x.z)));
        lu1 = T(log(u01(x.w)));
    }
};

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

struct WarpSmem {          // pointers into this warp's shared-memory slice
    double* th;
    double* best;
    double* mean;
    double* m2;
    double* comom;
    float* asf;
    float* psd;
    int* acc;
    int* rej;
    int* acct;
    double* sc;                  // [0] best_logp, [1] sum_lnL, [2] sum_lnL2
    unsigned long long* cnt;     // [0] shape evaluations, [1] far-path pixels, [2] likelihood passes
};

// ------------------------------------------------------------------------------------------------
// Row-streaming chain kernel.
//
// First version of this kernel kept the lane's pixels in registers and fully unrolled the
// per-row Voigt evaluation for PPL = 1/2/4/8: 34 k SASS instructions (544 KB) per kernel, and ncu
// showed the warps stalled on instruction fetch (stall_no_instruction 7.4 warps per issue,
// issue-active 32 %, profiles/r1_k2_a_unrolled_baseline.txt).  This version streams the rows
// (32 pixels each, two rows in flight for ILP) through ONE copy of the core / far code:
//   * du, flux (and 1/noise) of a row come from global memory - 3 KB per region, always L1 hits;
//   * the per-component shape cache, the tau of the other components ("base") and the
//     proposed shape live in warp-private shared memory, lane-private slots (no syncs);
//   * accepting a centroid/width step flips a slot index instead of copying the shape.
// The hot loop is now a few hundred instructions and the register footprint is halved.

template <typename T>
__device__ __forceinline__ T real_log(double x);
template <> __device__ __forceinline__ float real_log<float>(double x) { return __logf((float)x); }
template <> __device__ __forceinline__ double real_log<double>(double x) { return log(x); }

// One likelihood pass over the region, two rows (64 pixels) per trip, nothing masked: the
// region arrays and the shared-memory rows are padded to 64 pixels with neutral values, and
// everything is in PAIR layout (element jj*32 + lane = pixels 64 jj + lane and 64 jj + 32 + lane),
// so the two rows travel as one packed value through loads, arithmetic and shared memory.
// FRESH: evaluate the component `cc` and store its shape to Hdst; else read the shape from Hsrc.
// `An` and `basebuf` are pre-multiplied by NEGK (-log2(e) in fp32) so m = ex2(fma(An, H, base)).
// Returns sum (f - m)^2 (WT: noise-weighted) of m = exp(-(tau_others + A*H)) [convolved with the LSF].
template <typename T, int KIND, bool WT, bool LSF, bool FRESH>
__device__ __forceinline__ double chain_pass(const CompC<T>& cc, T An, const P2<T>* __restrict__ Hsrc,
                                             P2<T>* __restrict__ Hdst, const P2<T>* __restrict__ basebuf,
                                             const typename RS<T>::Pair* __restrict__ gdu,
                                             const typename RS<T>::Pair* __restrict__ gf,
                                             const typename RS<T>::Pair* __restrict__ gw, int npairs, int P, int lane,
                                             T* lsf_buf, const T* lsf_s, int taps, unsigned& nfar) {
    P2<T> S2 = p2_bc<T>(T(0));
    const P2<T> A2 = p2_bc<T>(An);
    int o = lane;                                   // pair-layout element of the current row pair
    T* pl = lsf_buf + (taps >> 1) + lane;
#pragma unroll 1
    for (int jj = 0; jj < npairs; ++jj, o += 32) {
        P2<T> H;
        if (FRESH) {
            H = pair_shape<T, KIND>(cc, p2_from(__ldg(gdu + o)), nfar);
            Hdst[o] = H;
        } else {
            H = Hsrc[o];
        }
        const P2<T> M = p2_exp_scaled<T>(p2_fma(A2, H, basebuf[o]));
        if (LSF) {
            pl[64 * jj] = p2_lo(M) - T(1);
            pl[64 * jj + 32] = p2_hi(M) - T(1);
        } else {
            P2<T> R = p2_add(p2_from(__ldg(gf + o)), p2_neg(M));
            if (WT) R = p2_mul(R, p2_from(__ldg(gw + o)));
            S2 = p2_fma(R, R, S2);
        }
    }
    T S = p2_lo(S2) + p2_hi(S2);
    if (LSF) {
        const int h = taps >> 1;
        for (int i = lane; i < 2 * h; i += 32) lsf_buf[(i < h) ? i : (64 * npairs + i)] = T(0);
        __syncwarp();
        const T* bq = lsf_buf + lane + 2 * h;
        const T* gfn = reinterpret_cast<const T*>(gf);
        const T* gwn = reinterpret_cast<const T*>(gw);
        for (int j = 0; j < 2 * npairs; ++j) {
            T acc = T(1);
            for (int t = 0; t < taps; ++t) acc = Real<T>::fma_(lsf_s[t], bq[-t], acc);
            // the convolution leaks into the padding pixels next to the region edge: mask them
            const int px = j * 32 + lane;
            const int e = 2 * (32 * (j >> 1) + lane) + (j & 1);        // natural pixel -> pair layout
            T rr = (px < P) ? (__ldg(gfn + e) - acc) : T(0);
            if (WT) rr *= __ldg(gwn + e);
            S = Real<T>::fma_(rr, rr, S);
            bq += 32;
        }
        __syncwarp();
    }
    return (double)warp_sum(S);
}

template <typename T, int KIND, bool WT, bool LSF>
__device__ void mcmc_chain(const McmcDev& p, int chain, int lane, const WarpSmem& ws, P2<T>* Hbuf, P2<T>* basebuf,
                           T* lsf_buf, const T* lsf_s, int HS) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const RegionSetDev& rs = p.rs;
    const int r = p.item_region[chain];
    const int P = rs.npix[r];
    const int n = p.item_ncomp[chain];
    const int ts = p.tstride;
    const int sdi = 4 * p.cmax;
    const int ncs = NP * n;                        // component steps per iteration
    const int k = ncs + (WT ? 0 : 1);
    const int taps = rs.lsf_taps;
    const int npairs = (P + 63) >> 6;
    const int tune_interval = p.tune_interval, thin = p.thin, burn = p.burn, tune_throughout = p.tune_throughout;
    const double ubase = rs.ubase[rs.tile_off[r]];
    typedef typename RS<T>::Pair Pair;
    const Pair* __restrict__ gdu = RS<T>::pdu(rs) + (rs.pix_off[r] >> 1);
    const Pair* __restrict__ gf = RS<T>::pf(rs) + (rs.pix_off[r] >> 1);
    const Pair* __restrict__ gw = RS<T>::pw(rs) + (rs.pix_off[r] >> 1);
    const T NK = Real<T>::NEGK();

    // ---- state -> shared
    const size_t gb = (size_t)chain * ts;
    for (int i = lane; i < ts; i += 32) {
        ws.th[i] = p.theta[gb + i];
        ws.asf[i] = p.asf[gb + i];
        ws.psd[i] = p.psd[gb + i];
        ws.acc[i] = p.acc[gb + i];
        ws.rej[i] = p.rej[gb + i];
        ws.acct[i] = p.acc_total[gb + i];
    }
    if (lane == 0) {
        ws.sc[0] = p.best_logp[chain];
        ws.sc[1] = p.sum_lnL[chain];
        ws.sc[2] = p.sum_lnL2[chain];
        ws.cnt[0] = ws.cnt[1] = ws.cnt[2] = 0ull;
    }
    int ntally = p.ntally[chain];
    __syncwarp();

    unsigned n_far = 0u, n_shape = 0u, n_like = 0u;      // per iteration; flushed to ws.cnt
    // slot table: component c's shape lives in Hbuf[slot(c)*HS ...]; slot n is the spare one
    unsigned long long slots = 0ull;
    for (int c = 0; c <= n; ++c) slots |= (unsigned long long)c << (4 * c);
    auto slot_of = [&](int c) -> int { return (int)((slots >> (4 * c)) & 15ull); };

    // ---- build the shape cache and the current sum of squares
    CompC<T> cz;
    cz.A = cz.cb = cz.g = cz.y = cz.y2 = cz.L2 = T(0);
    for (int jj = 0; jj < npairs; ++jj) basebuf[jj * 32 + lane] = p2_bc<T>(T(0));
    {
        unsigned dummy = 0u;
        for (int c = 0; c < n; ++c) {
            const CompC<T> cc = make_comp<T, KIND>(ws.th[VAMP_TH(c, 0)], ws.th[VAMP_TH(c, 1)], ws.th[VAMP_TH(c, 2)],
                                                   ws.th[VAMP_TH(c, 3)], ubase);
            P2<T>* Hd = Hbuf + slot_of(c) * HS;
            const P2<T> A2 = p2_bc<T>(NK * cc.A);
            for (int jj = 0; jj < npairs; ++jj) {
                const int o = jj * 32 + lane;
                const P2<T> H = pair_shape<T, KIND>(cc, p2_from(__ldg(gdu + o)), dummy);
                Hd[o] = H;
                basebuf[o] = p2_fma(A2, H, basebuf[o]);
            }
        }
        P2<T>* Hspare = Hbuf + slot_of(n) * HS;
        for (int jj = 0; jj < npairs; ++jj) Hspare[jj * 32 + lane] = p2_bc<T>(T(0));
        // base holds the full tau here: one cached pass with A = 0 gives the current S
    }
    unsigned nf32 = 0u;
    double Scur = chain_pass<T, KIND, WT, LSF, false>(cz, T(0), Hbuf + slot_of(n) * HS, nullptr, basebuf, gdu, gf, gw,
                                                      npairs, P, lane, lsf_buf, lsf_s, taps, nf32);
    double sd = WT ? 1.0 : ws.th[sdi];
    double logsd = WT ? 0.0 : log(sd);
    double half_inv_var = WT ? 0.5 : 0.5 / (sd * sd);
    double lp = 0.0;                                          // sum_c (ln A_c - A_c), kept incrementally
    for (int c = 0; c < n; ++c) {
        const double a = ws.th[VAMP_TH(c, 0)];
        lp += log(a) - a;
    }

    WarpRng<T> rng;
    {
        const unsigned long long uid = p.chain_uid[chain];
        rng.uid_lo = (uint32_t)uid;
        rng.uid_hi = (uint32_t)(uid >> 32);
        rng.k0 = (uint32_t)p.seed;
        rng.k1 = (uint32_t)(p.seed >> 32);
    }
    // step counter q = 64 * qblk + qslot; a (normal, log-uniform) pair per step
    unsigned long long qblk = ((unsigned long long)p.it0 * (unsigned long long)k) >> 6;
    int qslot = (int)(((unsigned long long)p.it0 * (unsigned long long)k) & 63ull);
    bool need_fill = true;                 // (re)fill at the first step and whenever a 64-step block starts

    for (int it = p.it0; it < p.it1; ++it) {
        // ---- tuning (PyMC 2: every tune_interval iterations; optionally only during burn-in)
        if (it > 0 && (it % tune_interval) == 0 && (tune_throughout || it <= burn)) {
            for (int i = lane; i < ts; i += 32) {
                const int a = ws.acc[i], b = ws.rej[i];
                if (a + b > 0) {
                    const float nf = ws.asf[i] * tune_factor((float)a / (float)(a + b));
                    if (nf > 0.0f) ws.asf[i] = nf;
                }
                ws.acc[i] = 0;
                ws.rej[i] = 0;
            }
            __syncwarp();
        }
        const bool post_burn = it >= burn;

        int c = 0, i = 0;
        double pA = 0.0, pc = 0.0, pw1 = 0.0, pw2 = 0.0;
#pragma unroll 1
        for (int st = 0; st < k; ++st) {
            // ---- one (normal, log-uniform) pair; the only RNG call site of the kernel
            if (need_fill) {
                rng.refill(qblk, lane);
                need_fill = false;
            }
            double z, lu;
            {
                const T zs = (qslot & 1) ? rng.z1 : rng.z0;
                const T ls = (qslot & 1) ? rng.lu1 : rng.lu0;
                z = (double)__shfl_sync(VAMP_FULL_MASK, zs, qslot >> 1);
                lu = (double)__shfl_sync(VAMP_FULL_MASK, ls, qslot >> 1);
            }
            if (++qslot == 64) {
                qslot = 0;
                ++qblk;
                need_fill = true;
            }
            bool accept = false;
            int idx;
            double prop;
            if (st < ncs) {
                idx = VAMP_TH(c, i);
                if (i == 0) {
                    // new component: base = tau of the others, rebuilt from the cache (no drift)
                    __syncwarp();
                    P2<T>* bb = basebuf + lane;
                    if (c == 0) {
                        // once per iteration: rebuild from the cache, so rounding never accumulates
                        for (int jj = 0; jj < npairs; ++jj) bb[jj * 32] = p2_bc<T>(T(0));
                        for (int c2 = 1; c2 < n; ++c2) {
                            const P2<T> A2 = p2_bc<T>(NK * T(ws.th[VAMP_TH(c2, 0)]));
                            const P2<T>* Hs = Hbuf + slot_of(c2) * HS + lane;
                            for (int jj = 0; jj < npairs; ++jj) bb[jj * 32] = p2_fma(A2, Hs[jj * 32], bb[jj * 32]);
                        }
                    } else {
                        // next component: base += tau_{c-1} - tau_c  (two FMAs per pixel instead of n - 1)
                        const P2<T> Ap = p2_bc<T>(NK * T(ws.th[VAMP_TH(c - 1, 0)])), Ac = p2_bc<T>(-NK * T(ws.th[idx]));
                        const P2<T>* Hp = Hbuf + slot_of(c - 1) * HS + lane;
                        const P2<T>* Hn = Hbuf + slot_of(c) * HS + lane;
                        for (int jj = 0; jj < npairs; ++jj)
                            bb[jj * 32] = p2_fma(Ac, Hn[jj * 32], p2_fma(Ap, Hp[jj * 32], bb[jj * 32]));
                    }
                    pA = ws.th[idx];
                    pc = ws.th[idx + 1];
                    pw1 = ws.th[idx + 2];
                    pw2 = ws.th[idx + 3];
                }
                const double cur = (i == 0) ? pA : (i == 1) ? pc : (i == 2) ? pw1 : pw2;
                prop = cur + z * (double)(ws.asf[idx] * ws.psd[idx]);
                const double c_hi = (double)(P - 1);
                const double w_hi = c_hi * (KIND == VAMP_KIND_VOIGT ? 0.5 * FWHM_PER_SIGMA : 0.5);
                bool ok;
                if (i == 0) ok = prop > 0.0;
                else if (i == 1) ok = (prop >= 0.0) && (prop <= c_hi);
                else if (KIND == VAMP_KIND_VOIGT && i == 2) ok = (prop >= 0.0) && (prop <= w_hi);
                else ok = (prop > 0.0) && (prop <= w_hi);
                if (ok) {
                    double dprior = 0.0, Snew;
                    const int sc = slot_of(c), sp = slot_of(n);
                    if (i == 0) {
                        Snew = chain_pass<T, KIND, WT, LSF, false>(cz, NK * T(prop), Hbuf + sc * HS, nullptr, basebuf, gdu,
                                                                   gf, gw, npairs, P, lane, lsf_buf, lsf_s, taps, nf32);
                        dprior = (double)real_log<T>((double)(T(prop) * Real<T>::rcp(T(pA)))) - (prop - pA);
                    } else {
                        const CompC<T> cc = make_comp_fast<T, KIND>(pA, (i == 1) ? prop : pc, (i == 2) ? prop : pw1,
                                                                    (i == 3) ? prop : pw2, ubase);
                        ++n_shape;
                        nf32 = 0u;
                        Snew = chain_pass<T, KIND, WT, LSF, true>(cc, NK * cc.A, nullptr, Hbuf + sp * HS, basebuf, gdu, gf,
                                                                  gw, npairs, P, lane, lsf_buf, lsf_s, taps, nf32);
                        n_far += nf32;
                    }
                    ++n_like;
                    accept = lu < (Scur - Snew) * half_inv_var + dprior;      // false for NaN
                    if (accept) {
                        Scur = Snew;
                        if (i == 0) {
                            pA = prop;
                            lp += dprior;
                        } else {
                            if (i == 1) pc = prop;
                            else if (i == 2) pw1 = prop;
                            else pw2 = prop;
                            // the proposed shape becomes the component's shape: swap slot indices
                            slots &= ~((15ull << (4 * c)) | (15ull << (4 * n)));
                            slots |= ((unsigned long long)sp << (4 * c)) | ((unsigned long long)sc << (4 * n));
                        }
                    }
                }
                if (++i == NP) {
                    i = 0;
                    ++c;
                }
            } else {
                idx = sdi;
                prop = sd + z * (double)(ws.asf[sdi] * ws.psd[sdi]);
                if (prop > 0.0 && prop <= 1.0) {
                    const double lprop = log(prop);
                    const double hv = 0.5 / (prop * prop);
                    accept = lu < -Scur * (hv - half_inv_var) - P * (lprop - logsd);
                    if (accept) {
                        sd = prop;
                        logsd = lprop;
                        half_inv_var = hv;
                    }
                }
            }
            {   // counters and state: every lane computes, lane 0 stores (no divergent branch)
                const int na = ws.acc[idx] + (accept ? 1 : 0);
                const int nr = ws.rej[idx] + (accept ? 0 : 1);
                const int nt = ws.acct[idx] + ((accept && post_burn) ? 1 : 0);
                const double nv = accept ? prop : ws.th[idx];
                if (lane == 0) {
                    ws.acc[idx] = na;
                    ws.rej[idx] = nr;
                    ws.acct[idx] = nt;
                    ws.th[idx] = nv;
                }
            }
        }
        __syncwarp();

        // ---- end of iteration: log-posterior of the current state
        const double lnL = WT ? (-0.5 * Scur + rs.lognorm[r])
                              : (-Scur * half_inv_var - P * (logsd + 0.91893853320467274178));
        const double logp = lnL + lp;
        const bool better = logp > ws.sc[0];
        const bool tally = post_burn && (it % thin) == 0;
        __syncwarp();
        if (lane == 0) {
            if (better) ws.sc[0] = logp;
            if (tally) {
                ws.sc[1] += lnL;
                ws.sc[2] += lnL * lnL;
            }
            ws.cnt[0] += n_shape;
            ws.cnt[1] += n_far - n_shape * (unsigned)(64 * npairs - P);   // padding pixels always count as far
            ws.cnt[2] += n_like;
        }
        n_shape = n_far = n_like = 0u;
        if (better) {
            for (int i2 = lane; i2 < ts; i2 += 32) ws.best[i2] = ws.th[i2];
        }
        if (tally) {
            ++ntally;
            const double inv = 1.0 / (double)ntally;
            const double wgt = (double)(ntally - 1) * inv;
            for (int i2 = lane; i2 < 6 * n; i2 += 32) {          // co-moments first (need the old means)
                const int cc2 = i2 / 6, pr = i2 - 6 * cc2;
                const int a = (pr < 3) ? 0 : (pr < 5 ? 1 : 2);
                const int b = (pr < 3) ? pr + 1 : (pr < 5 ? pr - 1 : 3);
                const double da = ws.th[VAMP_TH(cc2, a)] - ws.mean[VAMP_TH(cc2, a)];
                const double db = ws.th[VAMP_TH(cc2, b)] - ws.mean[VAMP_TH(cc2, b)];
                ws.comom[i2] += wgt * da * db;
            }
            __syncwarp();
            for (int i2 = lane; i2 < ts; i2 += 32) {
                const double x = ws.th[i2];
                const double d = x - ws.mean[i2];
                const double mn = ws.mean[i2] + d * inv;
                ws.mean[i2] = mn;
                ws.m2[i2] += d * (x - mn);
            }
            if (p.trace) {
                const int slot = ntally - 1;
                if (slot < p.trace_cap)
                    for (int i2 = lane; i2 < ts; i2 += 32)
                        p.trace[((size_t)chain * p.trace_cap + slot) * ts + i2] = ws.th[i2];
            }
        }
        __syncwarp();
    }

    // ---- state -> global
    for (int i = lane; i < ts; i += 32) {
        p.theta[gb + i] = ws.th[i];
        p.asf[gb + i] = ws.asf[i];
        p.acc[gb + i] = ws.acc[i];
        p.rej[gb + i] = ws.rej[i];
        p.acc_total[gb + i] = ws.acct[i];
    }
    if (lane == 0) {
        const double lnL = WT ? lnL_noise(Scur, rs.lognorm[r]) : lnL_sd(Scur, sd, P);
        double lpx = 0.0;
        for (int c = 0; c < n; ++c) {
            const double a = ws.th[VAMP_TH(c, 0)];
            lpx += log(a) - a;
        }
        p.lnL[chain] = lnL;
        p.logp[chain] = lnL + lpx;
        p.best_logp[chain] = ws.sc[0];
        p.ntally[chain] = ntally;
        p.sum_lnL[chain] = ws.sc[1];
        p.sum_lnL2[chain] = ws.sc[2];
        if (p.work) {
            p.work[3 * (size_t)chain + 0] += ws.cnt[0];    // component evaluations (x P pixels each)
            p.work[3 * (size_t)chain + 1] += ws.cnt[1];    // pixels of far-path rows (padding rows excluded below)
            p.work[3 * (size_t)chain + 2] += ws.cnt[2];    // exp(-tau) + residual passes (x P pixels)
        }
    }
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T>
__host__ __device__ inline size_t k2_warp_bytes(int ppl_max, int ncomp_max, int cmax, bool lsf) {
    const int ts = 4 * cmax + 1;
    size_t b = 0;
    b += align16(sizeof(double) * (ts + 6));                  // th, sc[3], cnt[3]
    b += align16(sizeof(float) * 2 * ts);                      // asf, psd
    b += align16(sizeof(int) * 3 * ts);                        // acc, rej, acct
    b += align16(sizeof(T) * (size_t)(ncomp_max + 2) * 32 * ppl_max); // shape cache (+ spare slot) + base
    if (lsf) b += align16(sizeof(T) * (32 * ppl_max + VAMP_MAX_LSF_TAPS));
    return b;
}

template <typename T, int KIND, bool WT, bool LSF>
__global__ void __launch_bounds__(K2_WARPS * 32, (sizeof(T) == 4) ? VAMP_K2_MINBLOCKS : 3)
k2_mcmc_kernel(McmcDev p, int ppl_max, int ncomp_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const size_t head = align16(sizeof(T) * VAMP_MAX_LSF_TAPS);
    if (LSF) {
        for (int i = threadIdx.x; i < p.rs.lsf_taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(p.rs)[i];
        __syncthreads();
    }
    const int chain = blockIdx.x * K2_WARPS + warp;
    if (chain >= p.B) return;
    const int ts = p.tstride;
    unsigned char* base = smem_raw + head + (size_t)warp * k2_warp_bytes<T>(ppl_max, ncomp_max, p.cmax, LSF);
    WarpSmem ws;
    ws.th = reinterpret_cast<double*>(base);
    ws.sc = ws.th + ts;
    ws.cnt = reinterpret_cast<unsigned long long*>(ws.sc + 3);
    base += align16(sizeof(double) * (ts + 6));
    // best state and tally statistics are touched rarely (improvements, every `thin` iterations):
    // they stay in global memory so that more chains fit on an SM
    ws.best = p.best_theta + (size_t)chain * ts;
    ws.mean = p.mean + (size_t)chain * ts;
    ws.m2 = p.m2 + (size_t)chain * ts;
    ws.comom = p.comom + (size_t)chain * 6 * p.cmax;
    ws.asf = reinterpret_cast<float*>(base);
    ws.psd = ws.asf + ts;
    base += align16(sizeof(float) * 2 * ts);
    ws.acc = reinterpret_cast<int*>(base);
    ws.rej = ws.acc + ts;
    ws.acct = ws.rej + ts;
    base += align16(sizeof(int) * 3 * ts);
    const int HS = 16 * ppl_max;                                // row pairs x 32 lanes, in P2 units
    P2<T>* Hbuf = reinterpret_cast<P2<T>*>(base);               // [(ncomp_max + 1)][HS]
    P2<T>* basebuf = Hbuf + (size_t)(ncomp_max + 1) * HS;       // [HS]
    base += align16(sizeof(T) * (size_t)(ncomp_max + 2) * 32 * ppl_max);
    T* lsf_buf = reinterpret_cast<T*>(base);
    mcmc_chain<T, KIND, WT, LSF>(p, chain, lane, ws, Hbuf, basebuf, lsf_buf, lsf_s, HS);
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
    p.best_logp[chain] = -INFINITY;
    p.logp[chain] = -INFINITY;
    p.lnL[chain] = -INFINITY;
    p.ntally[chain] = 0;
    p.sum_lnL[chain] = 0.0;
    p.sum_lnL2[chain] = 0.0;
}

template <typename T, int KIND, bool WT, bool LSF>
cudaError_t launch_k2_t(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    const int ppl_max = 2 * ((max_npix + 63) / 64);          // rows, in pairs of 32-pixel rows
    const size_t smem = align16(sizeof(T) * VAMP_MAX_LSF_TAPS) + K2_WARPS * k2_warp_bytes<T>(ppl_max, max_ncomp, p.cmax, LSF);
    auto kern = k2_mcmc_kernel<T, KIND, WT, LSF>;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (p.B + K2_WARPS - 1) / K2_WARPS;
    kern<<<grid, K2_WARPS * 32, smem, st>>>(p, ppl_max, max_ncomp);
    return cudaGetLastError();
}

}  // namespace

#if VAMP_INST_F64 == 0
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

cudaError_t VAMP_INST_NAME(vamp_launch_k2)(const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    if (p.B <= 0 || p.it1 <= p.it0) return cudaSuccess;
    if (max_npix > VAMP_TILE_PIX) return cudaErrorInvalidValue;
    const bool lsf = p.rs.lsf_taps > 1;
    if (p.sd_free)
        return lsf ? launch_k2_t<InstT, InstKind, false, true>(p, max_npix, max_ncomp, st)
                   : launch_k2_t<InstT, InstKind, false, false>(p, max_npix, max_ncomp, st);
    return lsf ? launch_k2_t<InstT, InstKind, true, true>(p, max_npix, max_ncomp, st)
               : launch_k2_t<InstT, InstKind, true, false>(p, max_npix, max_ncomp, st);
}
