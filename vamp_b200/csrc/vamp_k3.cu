// vamp_k3.cu - K3: batched Nelder-Mead MAP polish + BIC, one CTA (8 warps) per start.
//
// Replaces  mc.MAP(model).fit(iterlim, tol=1e-3)  and  map.BIC  (vamp_1.0/vpfits.py:357-358,
// 419-427): PyMC 2 minimises -logp (priors + Normal likelihood) with scipy.optimize.fmin
// (downhill simplex: rho=1, chi=2, psi=0.5, sigma=0.5; stop when max|x_i - x_0| <= xtol and
// max|f_i - f_0| <= ftol, or at maxiter / maxfun) from the last MCMC state, then reports
// BIC = k ln N - 2 lnL with the *likelihood* at the optimum (restated in oracle/sampler.py).
//
// Nelder-Mead is a serial chain of objective evaluations, so the latency of ONE evaluation is
// what matters: each of the CTA's 8 warps evaluates one 32-pixel row of the model (all
// components, same row_shape code as K2), the partial sums are combined through shared memory
// in a fixed order (bit-reproducible), and every thread replays the identical simplex logic on
// the shared simplex.  The first version ran a whole evaluation in one warp (8 rows in
// sequence); with 96 starts per bench step that was a 78 ms serial tail per fit.
#include "vamp_device.cuh"
#include "vamp_kernels.h"

namespace {

constexpr int K3_WARPS = 8;
constexpr int K3_THREADS = K3_WARPS * 32;
constexpr double FWHM_PER_SIGMA = 2.35482004503094938202;

__host__ __device__ inline size_t k3_align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T>
__host__ __device__ inline size_t k3_smem_bytes(int kmax, int ts, bool lsf, int ncmax, int max_npix = VAMP_TILE_PIX) {
    size_t b = 0;
    b += k3_align16(sizeof(double) * ((size_t)(kmax + 1) * ts + (kmax + 1) + 4 * ts + 4 * K3_WARPS));
    b += k3_align16(sizeof(CompC<T>) * K3_WARPS * (size_t)ncmax);
    if (lsf) b += k3_align16(sizeof(T) * (((size_t)max_npix + 63) / 64 * 64 + 2 * VAMP_MAX_LSF_TAPS));
    return b;
}

template <typename T, int KIND>
struct Objective {
    const MapDev& p;
    const RegionSetDev& rs;
    int P, n, sdi, taps, lane, warp, nrows;
    int nl;                // lines per evaluation: n, or 2 n in a multiplet region
    const double* mrow;    // the region's row of the multiplet table, or nullptr
    double ubase, lognorm, c_hi, w_hi;
    const double* tub;     // bases of the region's coordinate tiles (tub[0] = ubase): rows of tile t are moved to tile 0's base
    const T* gdu;
    const T* gf;
    const T* gw;
    double* red;           // [4 * K3_WARPS] shared
    T* mbuf;               // LSF staging (shared)
    CompC<T>* ccw;         // [K3_WARPS][ncmax] per-warp component constants (shared)
    int ncmax;
    const T* lsf_s;
    double lnL, chi2;      // of the last evaluation

    // -logp of theta (full layout, shared memory); +inf outside the prior support.
    // Called by ALL threads of the CTA with the same theta; returns the same value everywhere.
    __device__ __noinline__ double operator()(const double* th) {
        constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
        // Bounds, prior and component constants: lane c of EVERY warp handles component c (c + 32, ...), so the
        // float64 log and the divisions of make_comp run once in parallel instead of n times in sequence
        // (they were a third of an evaluation's latency); each warp stages its constants in its own slice.
        bool ok = true;
        double lp = 0.0;
        CompC<T>* ccs = ccw + warp * ncmax;
        for (int c = lane; c < n; c += 32) {
            const double a = th[VAMP_TH(c, 0)], ce = th[VAMP_TH(c, 1)], w1 = th[VAMP_TH(c, 2)], w2 = th[VAMP_TH(c, 3)];
            ok = ok && (a > 0.0) && (ce >= 0.0) && (ce <= c_hi);
            if (NP == 4) ok = ok && (w1 >= 0.0) && (w1 <= w_hi) && (w2 > 0.0) && (w2 <= w_hi);
            else ok = ok && (w1 > 0.0) && (w1 <= w_hi);
            lp += log(a) - a;
            ccs[c] = make_comp_k1<T, KIND, true>(a, ce, w1, w2, ubase);      // clamped 1/G: the padding coordinate stays finite
            if (nl > n) {                                                    // multiplet region: the component's second line
                double a2, ce2, w12, w22;
                MultD md;
                md.fr = mrow[0]; md.sh = mrow[1]; md.ar = mrow[2]; md.lr = mrow[3];
                line_params<KIND>(th, n + c, n, md, a2, ce2, w12, w22);
                ccs[n + c] = make_comp_k1<T, KIND, true>(a2, ce2, w12, w22, ubase);
            }
        }
        double sd = 1.0;
        if (p.sd_free) {
            sd = th[sdi];
            ok = ok && (sd > 0.0) && (sd <= 1.0);
        }
        ok = __all_sync(VAMP_FULL_MASK, ok);
        if (!ok) { lnL = -INFINITY; chi2 = INFINITY; return INFINITY; }
        lp = warp_sum(lp);
        __syncwarp();
        const int h = taps >> 1;
        T S = T(0), X = T(0);
        unsigned dummy = 0u;
        for (int row = warp; row < nrows; row += K3_WARPS) {
            const int px = row * 32 + lane;
            T du = __ldg(gdu + px);                       // rows are padded with neutral pixels
            if (nrows > VAMP_TILE_PIX / 32) du = T((double)du + (tub[row / (VAMP_TILE_PIX / 32)] - ubase));   // regions of several tiles
            T tau = T(0);
            // two components at a time as the halves of a two-wide value (two independent dependency chains: this
            // kernel is bound by the latency of one evaluation); one path per component pair and row
            for (int c = 0; c < nl; c += 2) {
                const CompC<T> c0 = ccs[c];
                CompC<T> c1 = c0;
                c1.A = T(0);
                if (c + 1 < nl) c1 = ccs[c + 1];
                const P2<T> X = p2_mul(p2_sub(p2_bc<T>(du), p2_make(c0.cb, c1.cb)), p2_make(c0.g, c1.g));
                P2<T> H;
                if (KIND == VAMP_KIND_GAUSS) {
                    const P2<T> Q = p2_mul(X, X);
                    H = p2_make(Real<T>::exp2_neg(p2_lo(Q)), Real<T>::exp2_neg(p2_hi(Q)));
                } else {
                    const P2<T> Y = p2_make(c0.y, c1.y);
                    const P2<T> S2 = p2_fma(X, X, p2_mul(Y, Y));
                    const bool far = __all_sync(VAMP_FULL_MASK, !(fmin(p2_lo(S2), p2_hi(S2)) < Fad<T>::R2()));
                    H = far ? fad_far2y<T>(X, Y, S2) : fad_core2y<T>(X, Y, S2);
                }
                tau = Real<T>::fma_(c0.A, p2_lo(H), Real<T>::fma_(c1.A, p2_hi(H), tau));
            }
            const T m = Real<T>::exp_neg(tau);
            if (taps > 1) {
                mbuf[h + px] = (px < P) ? (m - T(1)) : T(0);
            } else {
                const T r = __ldg(gf + px) - m;
                S = Real<T>::fma_(r, r, S);
                const T rw = r * __ldg(gw + px);
                X = Real<T>::fma_(rw, rw, X);
            }
        }
        if (taps > 1) {
            for (int i = threadIdx.x; i < 2 * h; i += K3_THREADS) mbuf[(i < h) ? i : (32 * nrows + i)] = T(0);
            __syncthreads();
            for (int row = warp; row < nrows; row += K3_WARPS) {
                const int px = row * 32 + lane;
                T acc = T(1);
                const T* bq = mbuf + px + 2 * h;
                for (int t = 0; t < taps; ++t) acc = Real<T>::fma_(lsf_s[t], bq[-t], acc);
                const T r = (px < P) ? (__ldg(gf + px) - acc) : T(0);
                S = Real<T>::fma_(r, r, S);
                const T rw = r * __ldg(gw + px);
                X = Real<T>::fma_(rw, rw, X);
            }
        }
        const double Sw = (double)warp_sum(S), Xw = (double)warp_sum(X);
        if (lane == 0) { red[warp] = Sw; red[K3_WARPS + warp] = Xw; }
        __syncthreads();
        double Sd = 0.0, Xd = 0.0;
#pragma unroll
        for (int w = 0; w < K3_WARPS; ++w) { Sd += red[w]; Xd += red[K3_WARPS + w]; }
        __syncthreads();                                  // red / mbuf may be rewritten by the next call
        chi2 = Xd;
        lnL = p.sd_free ? lnL_sd(Sd, sd, P) : lnL_noise(Xd, lognorm);
        const double f = -(lnL + lp);
        return (f == f) ? f : INFINITY;
    }
};

template <typename T, int KIND>
__device__ void nm_run(const MapDev& p, int s, double* sm, CompC<T>* ccw, int ncmax, T* mbuf, const T* lsf_s) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const RegionSetDev& rs = p.rs;
    const int ts = p.tstride;
    const int tid = threadIdx.x;
    const int r = p.item_region[s];
    Objective<T, KIND> f{p, rs};
    f.P = rs.npix[r];
    f.n = p.item_ncomp[s];
    f.mrow = region_mult_row(rs, r);
    f.nl = f.mrow ? 2 * f.n : f.n;
    f.sdi = 4 * p.cmax;
    f.taps = rs.lsf_taps;
    f.lane = tid & 31;
    f.warp = tid >> 5;
    f.nrows = 2 * ((f.P + 63) >> 6);
    f.ubase = rs.ubase[rs.tile_off[r]];
    f.tub = rs.ubase + rs.tile_off[r];
    f.lognorm = rs.lognorm[r];
    f.c_hi = (double)(f.P - 1);
    f.w_hi = 0.5 * (double)(f.P - 1) * (KIND == VAMP_KIND_VOIGT ? FWHM_PER_SIGMA : 1.0);
    f.gdu = RS<T>::du(rs) + rs.pix_off[r];
    f.gf = RS<T>::f(rs) + rs.pix_off[r];
    f.gw = RS<T>::w(rs) + rs.pix_off[r];
    f.mbuf = mbuf;
    f.ccw = ccw;
    f.ncmax = ncmax;
    f.lsf_s = lsf_s;

    const int n = f.n;
    const int k = NP * n + (p.sd_free ? 1 : 0);      // free dimension N of scipy
    const int nv = k + 1;
    auto is_free = [&](int i) -> bool {
        if (i == f.sdi) return p.sd_free != 0;
        return (i < 4 * n) && ((i & 3) < NP);
    };
    auto free_pos = [&](int j) -> int {               // position of the j-th free parameter
        if (j == NP * n) return f.sdi;
        return 4 * (j / NP) + (j % NP);
    };

    double* sim = sm;                       // [nv][ts]
    double* fsim = sim + (size_t)nv * ts;   // [nv]
    double* xbar = fsim + nv;               // [ts]
    double* xt = xbar + ts;                 // [ts] trial point
    double* xr = xt + ts;                   // [ts]
    double* x0 = xr + ts;                   // [ts]
    f.red = x0 + ts;                        // [4 * K3_WARPS]

    for (int i = tid; i < ts; i += K3_THREADS) x0[i] = p.theta[(size_t)s * ts + i];
    __syncthreads();
    // ---- initial simplex (scipy: vertex j+1 perturbs coordinate j by 5 %, or 0.00025 if zero)
    for (int v = 0; v < nv; ++v) {
        const int pj = v > 0 ? free_pos(v - 1) : -1;
        for (int i = tid; i < ts; i += K3_THREADS) {
            double x = x0[i];
            if (i == pj) x = (x != 0.0) ? 1.05 * x : 0.00025;
            sim[(size_t)v * ts + i] = x;
        }
    }
    __syncthreads();
    int nfev = 0;
    for (int v = 0; v < nv; ++v) {
        const double fv = f(sim + (size_t)v * ts);
        ++nfev;
        if (tid == 0) fsim[v] = fv;
    }
    __syncthreads();

    int iters = 1;
    int ib = 0, iw = 0, iw2 = 0;
    double fb = 0, fw = 0, fw2 = 0;
    const int maxfun = p.maxfun > 0 ? p.maxfun : 0x7fffffff;
    while (true) {
        // ---- order statistics in one pass: best (first minimum), worst and second worst (last maxima,
        //      the order a stable sort gives); tiny (<= 34 values), identical in every thread
        fb = INFINITY; fw = -INFINITY; fw2 = -INFINITY; ib = iw = iw2 = 0;
        for (int v = 0; v < nv; ++v) {
            const double x = fsim[v];
            if (x < fb) { fb = x; ib = v; }
            if (x >= fw) { fw2 = fw; iw2 = iw; fw = x; iw = v; }
            else if (x >= fw2) { fw2 = x; iw2 = v; }
        }
        if (!(nfev < maxfun && iters < p.maxiter)) break;
        // ---- convergence test (identical in every thread: all read the same shared simplex).  Both
        //      conditions must hold, so the O(k nv) coordinate spread is only formed once the function
        //      spread is within ftol.
        double df = 0.0;
        for (int v = 0; v < nv; ++v) df = fmax(df, fabs(fsim[v] - fb));
        if (df <= p.ftol) {
            double dx = 0.0;
            for (int i = 0; i < ts; ++i) {
                if (!is_free(i)) continue;
                const double b = sim[(size_t)ib * ts + i];
                for (int v = 0; v < nv; ++v) dx = fmax(dx, fabs(sim[(size_t)v * ts + i] - b));
                if (dx > p.xtol) break;
            }
            if (dx <= p.xtol) break;
        }

        // ---- centroid of all but the worst
        for (int i = tid; i < ts; i += K3_THREADS) {
            double a = 0.0;
            for (int v = 0; v < nv; ++v)
                if (v != iw) a += sim[(size_t)v * ts + i];
            xbar[i] = is_free(i) ? a / (double)k : sim[(size_t)ib * ts + i];
        }
        __syncthreads();
        const double* xw = sim + (size_t)iw * ts;
        auto combine = [&](double* dst, double ca, double cb_) {     // dst = ca*xbar + cb*xw
            for (int i = tid; i < ts; i += K3_THREADS) dst[i] = is_free(i) ? ca * xbar[i] + cb_ * xw[i] : xbar[i];
            __syncthreads();
        };
        auto replace_worst = [&](const double* src, double fv) {
            for (int i = tid; i < ts; i += K3_THREADS) sim[(size_t)iw * ts + i] = src[i];
            if (tid == 0) fsim[iw] = fv;
            __syncthreads();
        };
        combine(xr, 2.0, -1.0);                                     // reflection (rho = 1)
        const double fxr = f(xr);
        ++nfev;
        bool shrink = false;
        if (fxr < fb) {
            combine(xt, 3.0, -2.0);                                 // expansion (1 + rho*chi, -rho*chi)
            const double fxe = f(xt);
            ++nfev;
            if (fxe < fxr) replace_worst(xt, fxe);
            else replace_worst(xr, fxr);
        } else if (fxr < fw2) {
            replace_worst(xr, fxr);
        } else if (fxr < fw) {
            combine(xt, 1.5, -0.5);                                 // outside contraction
            const double fxc = f(xt);
            ++nfev;
            if (fxc <= fxr) replace_worst(xt, fxc);
            else shrink = true;
        } else {
            combine(xt, 0.5, 0.5);                                  // inside contraction
            const double fxcc = f(xt);
            ++nfev;
            if (fxcc < fw) replace_worst(xt, fxcc);
            else shrink = true;
        }
        if (shrink) {
            for (int v = 0; v < nv; ++v) {
                if (v == ib) continue;
                for (int i = tid; i < ts; i += K3_THREADS) {
                    const double b = sim[(size_t)ib * ts + i];
                    sim[(size_t)v * ts + i] = b + 0.5 * (sim[(size_t)v * ts + i] - b);
                }
                __syncthreads();
                const double fv = f(sim + (size_t)v * ts);
                ++nfev;
                if (tid == 0) fsim[v] = fv;
            }
            __syncthreads();
        }
        ++iters;
    }
    // ---- outputs at the best vertex
    const double fbest = f(sim + (size_t)ib * ts);
    for (int i = tid; i < ts; i += K3_THREADS) p.theta[(size_t)s * ts + i] = sim[(size_t)ib * ts + i];
    if (tid == 0) {
        p.out_logp[s] = -fbest;
        p.out_lnL[s] = f.lnL;
        p.out_chi2[s] = f.chi2;
        p.out_bic[s] = (double)k * log((double)f.P) - 2.0 * f.lnL;
        p.out_iters[s] = iters;
        p.out_nfev[s] = nfev;
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(K3_THREADS)
k3_map_kernel(MapDev p, int kmax, int ncmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool LSF = p.rs.lsf_taps > 1;
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const size_t head = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS);
    if (LSF) {
        for (int i = threadIdx.x; i < p.rs.lsf_taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(p.rs)[i];
    }
    __syncthreads();
    const int s = p.S - 1 - (int)blockIdx.x;           // the host sorts starts by ascending cost: longest first
    if (s < 0) return;
    unsigned char* base = smem_raw + head;
    double* sm = reinterpret_cast<double*>(base);
    base += k3_align16(sizeof(double) * ((size_t)(kmax + 1) * p.tstride + (kmax + 1) + 4 * p.tstride + 4 * K3_WARPS));
    CompC<T>* ccw = reinterpret_cast<CompC<T>*>(base);
    base += k3_align16(sizeof(CompC<T>) * K3_WARPS * (size_t)ncmax);
    T* mbuf = reinterpret_cast<T*>(base);
    nm_run<T, KIND>(p, s, sm, ccw, ncmax, mbuf, lsf_s);
}


// ================================================================================================
// K3w - the throughput schedule: ONE WARP per start, objective in fp32 (packed two-row Faddeeva,
// same code as K2), simplex and Nelder-Mead logic in float64, and the reported lnL / chi^2 / BIC
// re-evaluated in float64 at the optimum.  With thousands of starts (10^5 regions x candidates
// x restarts) the CTA-per-start schedule above leaves most of an SM waiting on one simplex; here
// every warp is busy and 16+ simplexes share an SM.  Lanes own pixels during an evaluation and
// coordinates during the simplex updates.
#if VAMP_INST_F64 == 0
constexpr int K3W_WARPS = 4;
#ifndef VAMP_K3W_SWITCH
#define VAMP_K3W_SWITCH 0.01
#endif
#ifndef VAMP_K3W_MINBLOCKS
#define VAMP_K3W_MINBLOCKS 4
#endif
constexpr double K3W_SWITCH = VAMP_K3W_SWITCH;     // function spread at which the search moves to the float64 objective

template <typename T>
struct CompK { T A, cb, g, y; };          // per-component constants staged in shared memory

__host__ __device__ inline size_t k3w_warp_bytes(int kmax, int ts, int ncomp_max) {
    return k3_align16(sizeof(double) * ((size_t)(kmax + 1) * ts + (kmax + 1) + 3 * ts)) +
           k3_align16(sizeof(CompK<double>) * ncomp_max);
}

// -logp of `th` (component slots + sd at index `sdi`, shared memory); +inf outside the prior support.  All lanes
// get the same value.  T = float: search precision; T = double: reporting precision.
template <typename T, int KIND, typename F = Fad<T>>
__device__ __noinline__ double k3w_objective(const MapDev& p, const double* th, int sdi, int r, int P, int n, int lane,
                                             void* ccbuf, double c_hi, double w_hi, double ubase, double& lnL, double& chi2) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    typedef typename RS<T>::Pair Pair;
    const RegionSetDev& rs = p.rs;
    CompK<T>* ccs = reinterpret_cast<CompK<T>*>(ccbuf);
    const double* mrow = region_mult_row(rs, r);              // the region's row of the multiplet table, or nullptr
    const int nl = mrow ? 2 * n : n;                          // lines per evaluation
    // ---- bounds, prior and component constants: lane c handles component c (c + 32, ...)
    bool ok = true;
    double lp = 0.0;
    for (int c = lane; c < n; c += 32) {
        const double a = th[VAMP_TH(c, 0)], ce = th[VAMP_TH(c, 1)], w1 = th[VAMP_TH(c, 2)], w2 = th[VAMP_TH(c, 3)];
        ok = ok && (a > 0.0) && (ce >= 0.0) && (ce <= c_hi);
        if (NP == 4) ok = ok && (w1 >= 0.0) && (w1 <= w_hi) && (w2 > 0.0) && (w2 <= w_hi);
        else ok = ok && (w1 > 0.0) && (w1 <= w_hi);
        lp += log(a) - a;
        const CompC<T> cc = make_comp_k1<T, KIND, true>(a, ce, w1, w2, ubase);
        ccs[c].A = Real<T>::NEGK() * cc.A;
        ccs[c].cb = cc.cb;
        ccs[c].g = cc.g;
        ccs[c].y = cc.y;
        if (nl > n) {                          // multiplet region: the component's second line
            double a2, ce2, w12, w22;
            MultD md;
            md.fr = mrow[0]; md.sh = mrow[1]; md.ar = mrow[2]; md.lr = mrow[3];
            line_params<KIND>(th, n + c, n, md, a2, ce2, w12, w22);
            const CompC<T> c2 = make_comp_k1<T, KIND, true>(a2, ce2, w12, w22, ubase);
            ccs[n + c].A = Real<T>::NEGK() * c2.A;
            ccs[n + c].cb = c2.cb;
            ccs[n + c].g = c2.g;
            ccs[n + c].y = c2.y;
        }
    }
    double sd = 1.0;
    if (p.sd_free) {
        sd = th[sdi];
        ok = ok && (sd > 0.0) && (sd <= 1.0);
    }
    ok = __all_sync(VAMP_FULL_MASK, ok);
    if (!ok) { lnL = -INFINITY; chi2 = INFINITY; return INFINITY; }
    lp = warp_sum(lp);
    __syncwarp();
    // ---- model and residuals, two rows per trip (pair layout of the region set)
    const int npairs = (P + 63) >> 6;
    const Pair* gdu = RS<T>::pdu(rs) + (rs.pix_off[r] >> 1) + lane;
    const Pair* gf = RS<T>::pf(rs) + (rs.pix_off[r] >> 1) + lane;
    const Pair* gw = RS<T>::pw(rs) + (rs.pix_off[r] >> 1) + lane;
    P2<T> S2 = p2_bc<T>(T(0)), X2 = p2_bc<T>(T(0));
    unsigned nfar = 0u;
#pragma unroll 1
    for (int e = 0; e < 32 * npairs; e += 32) {
        const P2<T> DU = p2_from(__ldg(gdu + e));
        P2<T> tau = p2_bc<T>(T(0));
#pragma unroll 1
        for (int c = 0; c < nl; ++c) {
            CompC<T> cc;
            cc.A = ccs[c].A;
            cc.cb = ccs[c].cb;
            cc.g = ccs[c].g;
            cc.y = ccs[c].y;
            cc.y2 = cc.y * cc.y;
            cc.L2 = T(0);
            P2<T> X, S;
            bool far;
            pair_classify<T, KIND, F>(cc, DU, X, S, far);
            tau = p2_fma(p2_bc<T>(cc.A), pair_eval<T, KIND, F>(cc, X, S, far, nfar), tau);
        }
        // padding pixels: du = 1e12 gives tau = 0 exactly, flux 1, weight 0 (vamp_regions_create)
        const P2<T> R = p2_sub(p2_from(__ldg(gf + e)), p2_exp_scaled<T>(tau));
        S2 = p2_fma(R, R, S2);
        const P2<T> RW = p2_mul(R, p2_from(__ldg(gw + e)));
        X2 = p2_fma(RW, RW, X2);
    }
    const double Sd = (double)warp_sum(p2_lo(S2) + p2_hi(S2)), Xd = (double)warp_sum(p2_lo(X2) + p2_hi(X2));
    chi2 = Xd;
    lnL = p.sd_free ? lnL_sd(Sd, sd, P) : lnL_noise(Xd, rs.lognorm[r]);
    const double f = -(lnL + lp);
    return (f == f) ? f : INFINITY;
}

template <int KIND>
__global__ void __launch_bounds__(K3W_WARPS * 32, VAMP_K3W_MINBLOCKS)
k3w_map_kernel(MapDev p, int kmax, int ncomp_max) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = p.S - 1 - (blockIdx.x * K3W_WARPS + warp);      // starts sorted by ascending cost: longest first
    if (s < 0) return;
    // A vertex in shared memory holds the component slots of THIS LAUNCH (4 * mc values, mc = the launch's largest
    // component count) and sd - not the batch-wide stride of the global parameter vectors: the simplex is
    // (k + 1) x vs doubles per warp, and its size decides how many warps an SM holds (the host launches the starts
    // in groups of similar component count for that reason, vamp_b200/batch.py).
    const int gts = p.tstride;                        // stride and sd index of the global parameter vectors
    const int gsd = 4 * p.cmax;
    const int mc = (kmax - 1) / NP;
    const int ts = 4 * mc + 1;                        // vertex stride in shared memory
    unsigned char* wbase = smem_raw + (size_t)warp * k3w_warp_bytes(kmax, ts, ncomp_max);
    const RegionSetDev& rs = p.rs;
    const int r = p.item_region[s];
    const int P = rs.npix[r];
    const int n = p.item_ncomp[s];
    const int sdi = 4 * mc;                           // sd inside a shared-memory vertex
    const double ubase = rs.ubase[rs.tile_off[r]];
    const double c_hi = (double)(P - 1);
    const double w_hi = 0.5 * (double)(P - 1) * (KIND == VAMP_KIND_VOIGT ? FWHM_PER_SIGMA : 1.0);
    const int k = NP * n + (p.sd_free ? 1 : 0);
    const int nv = k + 1;
    auto is_free = [&](int i) -> bool {
        if (i == sdi) return p.sd_free != 0;
        return (i < 4 * n) && ((i & 3) < NP);
    };
    auto free_pos = [&](int j) -> int { return (j == NP * n) ? sdi : 4 * (j / NP) + (j % NP); };

    double* sim = reinterpret_cast<double*>(wbase);   // [nv][ts]
    double* fsim = sim + (size_t)nv * ts;              // [nv]
    double* xbar = fsim + nv;                          // [ts]
    double* xt = xbar + ts;                            // [ts]
    double* xr = xt + ts;                              // [ts]
    void* ccbuf = wbase + k3_align16(sizeof(double) * ((size_t)(kmax + 1) * ts + (kmax + 1) + 3 * ts));
    double lnL = 0.0, chi2 = 0.0;
    // Two stages: the fp32 objective differs from float64 by 1e-4 .. 1e-3 in logp, pixel-to-pixel jumps up
    // to 2e-3 (measured, tools/noise_scan.py) - the size of the stopping tolerance - so once the simplex's
    // function spread is within K3W_SWITCH the search continues on the float64 objective (all vertices
    // re-evaluated).  Before that every comparison is far above the noise and the two objectives decide alike.
    // The float64 stage of the SEARCH uses the mid-fidelity set (FadMid: |dH| <= 2.5e-10, no rounding noise); what is
    // reported at the end is evaluated with the full float64 set.
    bool fine = false, report = false;
    auto F = [&](const double* th) -> double {
        if (report) return k3w_objective<double, KIND>(p, th, sdi, r, P, n, lane, ccbuf, c_hi, w_hi, ubase, lnL, chi2);
        return fine ? k3w_objective<double, KIND, FadMid>(p, th, sdi, r, P, n, lane, ccbuf, c_hi, w_hi, ubase, lnL, chi2)
                    : k3w_objective<float, KIND>(p, th, sdi, r, P, n, lane, ccbuf, c_hi, w_hi, ubase, lnL, chi2);
    };

    // ---- initial simplex (scipy: vertex j+1 perturbs coordinate j by 5 %, or 0.00025 if zero)
    for (int v = 0; v < nv; ++v) {
        const int pj = v > 0 ? free_pos(v - 1) : -1;
        for (int i = lane; i < ts; i += 32) {
            double x = p.theta[(size_t)s * gts + ((i == sdi) ? gsd : i)];
            if (i == pj) x = (x != 0.0) ? 1.05 * x : 0.00025;
            sim[(size_t)v * ts + i] = x;
        }
    }
    __syncwarp();
    int nfev = 0;
    for (int v = 0; v < nv; ++v) {
        const double fv = F(sim + (size_t)v * ts);
        ++nfev;
        if (lane == 0) fsim[v] = fv;
    }
    __syncwarp();

    int iters = 1, ib = 0, iw = 0, iw2 = 0;
    double fb = 0, fw = 0, fw2 = 0;
    const int maxfun = p.maxfun > 0 ? p.maxfun : 0x7fffffff;
    while (true) {
        fb = INFINITY; fw = -INFINITY; fw2 = -INFINITY; ib = iw = iw2 = 0;
        for (int v = 0; v < nv; ++v) {
            const double x = fsim[v];
            if (x < fb) { fb = x; ib = v; }
            if (x >= fw) { fw2 = fw; iw2 = iw; fw = x; iw = v; }
            else if (x >= fw2) { fw2 = x; iw2 = v; }
        }
        if (!(nfev < maxfun && iters < p.maxiter)) break;
        double df = 0.0;
        for (int v = 0; v < nv; ++v) df = fmax(df, fabs(fsim[v] - fb));
        if (!fine && df <= K3W_SWITCH) {
            fine = true;
            for (int v = 0; v < nv; ++v) {
                const double fv = F(sim + (size_t)v * ts);
                ++nfev;
                if (lane == 0) fsim[v] = fv;
            }
            __syncwarp();
            continue;                              // order statistics again, on float64 values
        }
        if (df <= p.ftol) {
            double dx = 0.0;                       // lane i: spread of coordinate i
            for (int i = lane; i < ts; i += 32) {
                if (!is_free(i)) continue;
                const double b = sim[(size_t)ib * ts + i];
                for (int v = 0; v < nv; ++v) dx = fmax(dx, fabs(sim[(size_t)v * ts + i] - b));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dx = fmax(dx, __shfl_xor_sync(VAMP_FULL_MASK, dx, o));
            if (dx <= p.xtol) break;
        }
        // ---- centroid of all but the worst (lane-parallel over coordinates)
        for (int i = lane; i < ts; i += 32) {
            double a = 0.0;
            for (int v = 0; v < nv; ++v)
                if (v != iw) a += sim[(size_t)v * ts + i];
            xbar[i] = is_free(i) ? a / (double)k : sim[(size_t)ib * ts + i];
        }
        __syncwarp();
        const double* xw = sim + (size_t)iw * ts;
        auto combine = [&](double* dst, double ca, double cb_) {
            for (int i = lane; i < ts; i += 32) dst[i] = is_free(i) ? ca * xbar[i] + cb_ * xw[i] : xbar[i];
            __syncwarp();
        };
        auto replace_worst = [&](const double* src, double fv) {
            for (int i = lane; i < ts; i += 32) sim[(size_t)iw * ts + i] = src[i];
            if (lane == 0) fsim[iw] = fv;
            __syncwarp();
        };
        combine(xr, 2.0, -1.0);
        const double fxr = F(xr);
        ++nfev;
        bool shrink = false;
        if (fxr < fb) {
            combine(xt, 3.0, -2.0);
            const double fxe = F(xt);
            ++nfev;
            if (fxe < fxr) replace_worst(xt, fxe);
            else replace_worst(xr, fxr);
        } else if (fxr < fw2) {
            replace_worst(xr, fxr);
        } else if (fxr < fw) {
            combine(xt, 1.5, -0.5);
            const double fxc = F(xt);
            ++nfev;
            if (fxc <= fxr) replace_worst(xt, fxc);
            else shrink = true;
        } else {
            combine(xt, 0.5, 0.5);
            const double fxcc = F(xt);
            ++nfev;
            if (fxcc < fw) replace_worst(xt, fxcc);
            else shrink = true;
        }
        if (shrink) {
            for (int v = 0; v < nv; ++v) {
                if (v == ib) continue;
                for (int i = lane; i < ts; i += 32) {
                    const double b = sim[(size_t)ib * ts + i];
                    sim[(size_t)v * ts + i] = b + 0.5 * (sim[(size_t)v * ts + i] - b);
                }
                __syncwarp();
                const double fv = F(sim + (size_t)v * ts);
                ++nfev;
                if (lane == 0) fsim[v] = fv;
            }
            __syncwarp();
        }
        ++iters;
    }
    // ---- report in float64 at the best vertex
    report = true;
    const double fbest = F(sim + (size_t)ib * ts);
    for (int i = lane; i < ts; i += 32) p.theta[(size_t)s * gts + ((i == sdi) ? gsd : i)] = sim[(size_t)ib * ts + i];
    if (lane == 0) {
        p.out_logp[s] = -fbest;
        p.out_lnL[s] = lnL;
        p.out_chi2[s] = chi2;
        p.out_bic[s] = (double)k * log((double)P) - 2.0 * lnL;
        p.out_iters[s] = iters;
        p.out_nfev[s] = nfev;
    }
}
#endif  // VAMP_INST_F64 == 0

}  // namespace

cudaError_t VAMP_INST_NAME(vamp_launch_k3)(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    typedef InstT T;
    if (p.S <= 0) return cudaSuccess;
    // (this schedule reads the pixels from global memory row by row: regions of any length, tile by tile)
    const bool LSF = p.rs.lsf_taps > 1;
    const int NP = (InstKind == VAMP_KIND_VOIGT) ? 4 : 3;
    const int kmax = NP * max_ncomp + 1;
    const int nl_max = max_ncomp * (p.rs.mult ? 2 : 1);                 // multiplet regions: two lines per component
    const size_t smem = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS) + k3_smem_bytes<T>(kmax, p.tstride, LSF, nl_max, max_npix);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = k3_map_kernel<T, InstKind>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.S, K3_THREADS, smem, st>>>(p, kmax, nl_max);
    return cudaGetLastError();
}

#if VAMP_INST_F64 == 0
// mixed-precision warp-per-start schedule (vamp_map_dev dtype VAMP_MIXED); LSF regions use the CTA schedule
cudaError_t VAMP_CAT3(vamp_launch_k3w, mixed, VAMP_KN)(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    if (p.S <= 0) return cudaSuccess;
    if (max_npix > VAMP_TILE_PIX || p.rs.lsf_taps > 1) return cudaErrorInvalidValue;
    const int NP = (InstKind == VAMP_KIND_VOIGT) ? 4 : 3;
    const int kmax = NP * max_ncomp + 1;
    const int nl_max = max_ncomp * (p.rs.mult ? 2 : 1);
    const size_t smem = K3W_WARPS * k3w_warp_bytes(kmax, 4 * max_ncomp + 1, nl_max);     // vertex stride 4 mc + 1
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = k3w_map_kernel<InstKind>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(p.S + K3W_WARPS - 1) / K3W_WARPS, K3W_WARPS * 32, smem, st>>>(p, kmax, nl_max);
    return cudaGetLastError();
}
#endif
