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
__host__ __device__ inline size_t k3_smem_bytes(int kmax, int ts, bool lsf) {
    size_t b = 0;
    b += k3_align16(sizeof(double) * ((size_t)(kmax + 1) * ts + (kmax + 1) + 4 * ts + 4 * K3_WARPS));
    if (lsf) b += k3_align16(sizeof(T) * (VAMP_TILE_PIX + 2 * VAMP_MAX_LSF_TAPS));
    return b;
}

template <typename T, int KIND>
struct Objective {
    const MapDev& p;
    const RegionSetDev& rs;
    int P, n, sdi, taps, lane, warp, nrows;
    double ubase, lognorm, c_hi, w_hi;
    const T* gdu;
    const T* gf;
    const T* gw;
    double* red;           // [4 * K3_WARPS] shared
    T* mbuf;               // LSF staging (shared)
    const T* lsf_s;
    double lnL, chi2;      // of the last evaluation

    // -logp of theta (full layout, shared memory); +inf outside the prior support.
    // Called by ALL threads of the CTA with the same theta; returns the same value everywhere.
    __device__ __noinline__ double operator()(const double* th) {
        constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
        bool ok = true;
        double lp = 0.0;
        for (int c = 0; c < n; ++c) {
            const double a = th[VAMP_TH(c, 0)], ce = th[VAMP_TH(c, 1)], w1 = th[VAMP_TH(c, 2)], w2 = th[VAMP_TH(c, 3)];
            ok = ok && (a > 0.0) && (ce >= 0.0) && (ce <= c_hi);
            if (NP == 4) ok = ok && (w1 >= 0.0) && (w1 <= w_hi) && (w2 > 0.0) && (w2 <= w_hi);
            else ok = ok && (w1 > 0.0) && (w1 <= w_hi);
            lp += log(a) - a;
        }
        double sd = 1.0;
        if (p.sd_free) {
            sd = th[sdi];
            ok = ok && (sd > 0.0) && (sd <= 1.0);
        }
        if (!ok) { lnL = -INFINITY; chi2 = INFINITY; return INFINITY; }
        const int h = taps >> 1;
        T S = T(0), X = T(0);
        unsigned dummy = 0u;
        for (int row = warp; row < nrows; row += K3_WARPS) {
            const int px = row * 32 + lane;
            const T du = __ldg(gdu + px);                 // rows are padded with neutral pixels
            T tau = T(0);
            for (int c = 0; c < n; ++c) {
                const CompC<T> cc = make_comp<T, KIND>(th[VAMP_TH(c, 0)], th[VAMP_TH(c, 1)], th[VAMP_TH(c, 2)],
                                                       th[VAMP_TH(c, 3)], ubase);
                tau = Real<T>::fma_(cc.A, row_shape<T, KIND>(cc, du, dummy), tau);
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
__device__ void nm_run(const MapDev& p, int s, double* sm, T* mbuf, const T* lsf_s) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const RegionSetDev& rs = p.rs;
    const int ts = p.tstride;
    const int tid = threadIdx.x;
    const int r = p.item_region[s];
    Objective<T, KIND> f{p, rs};
    f.P = rs.npix[r];
    f.n = p.item_ncomp[s];
    f.sdi = 4 * p.cmax;
    f.taps = rs.lsf_taps;
    f.lane = tid & 31;
    f.warp = tid >> 5;
    f.nrows = 2 * ((f.P + 63) >> 6);
    f.ubase = rs.ubase[rs.tile_off[r]];
    f.lognorm = rs.lognorm[r];
    f.c_hi = (double)(f.P - 1);
    f.w_hi = 0.5 * (double)(f.P - 1) * (KIND == VAMP_KIND_VOIGT ? FWHM_PER_SIGMA : 1.0);
    f.gdu = RS<T>::du(rs) + rs.pix_off[r];
    f.gf = RS<T>::f(rs) + rs.pix_off[r];
    f.gw = RS<T>::w(rs) + rs.pix_off[r];
    f.mbuf = mbuf;
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
        // ---- order statistics: best, worst, second worst (ties as a stable sort would order them)
        fb = INFINITY; fw = -INFINITY; fw2 = -INFINITY; ib = iw = iw2 = 0;
        for (int v = 0; v < nv; ++v) {          // tiny (<= 34), identical in every thread
            const double x = fsim[v];
            if (x < fb) { fb = x; ib = v; }
        }
        for (int v = 0; v < nv; ++v) {
            const double x = fsim[v];
            if (x >= fw) { fw = x; iw = v; }
        }
        for (int v = 0; v < nv; ++v) {
            if (v == iw) continue;
            const double x = fsim[v];
            if (x >= fw2) { fw2 = x; iw2 = v; }
        }
        if (!(nfev < maxfun && iters < p.maxiter)) break;
        // ---- convergence test (identical in every thread: all read the same shared simplex)
        double dx = 0.0;
        for (int i = 0; i < ts; ++i) {
            if (!is_free(i)) continue;
            const double b = sim[(size_t)ib * ts + i];
            for (int v = 0; v < nv; ++v) dx = fmax(dx, fabs(sim[(size_t)v * ts + i] - b));
            if (dx > p.xtol) break;
        }
        double df = 0.0;
        for (int v = 0; v < nv; ++v) df = fmax(df, fabs(fsim[v] - fb));
        if (dx <= p.xtol && df <= p.ftol) break;

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
k3_map_kernel(MapDev p, int kmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool LSF = p.rs.lsf_taps > 1;
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const size_t head = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS);
    if (LSF) {
        for (int i = threadIdx.x; i < p.rs.lsf_taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(p.rs)[i];
    }
    __syncthreads();
    const int s = blockIdx.x;
    if (s >= p.S) return;
    unsigned char* base = smem_raw + head;
    double* sm = reinterpret_cast<double*>(base);
    T* mbuf = reinterpret_cast<T*>(base + k3_align16(sizeof(double) * ((size_t)(kmax + 1) * p.tstride + (kmax + 1) +
                                                                        4 * p.tstride + 4 * K3_WARPS)));
    nm_run<T, KIND>(p, s, sm, mbuf, lsf_s);
}

}  // namespace

cudaError_t VAMP_INST_NAME(vamp_launch_k3)(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    typedef InstT T;
    if (p.S <= 0) return cudaSuccess;
    if (max_npix > VAMP_TILE_PIX) return cudaErrorInvalidValue;
    const bool LSF = p.rs.lsf_taps > 1;
    const int NP = (InstKind == VAMP_KIND_VOIGT) ? 4 : 3;
    const int kmax = NP * max_ncomp + 1;
    const size_t smem = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS) + k3_smem_bytes<T>(kmax, p.tstride, LSF);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = k3_map_kernel<T, InstKind>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.S, K3_THREADS, smem, st>>>(p, kmax);
    return cudaGetLastError();
}
