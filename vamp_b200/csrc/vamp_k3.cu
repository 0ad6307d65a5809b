// vamp_k3.cu - K3: batched Nelder-Mead MAP polish + BIC, one warp per start.
//
// Replaces  mc.MAP(model).fit(iterlim, tol=1e-3)  and  map.BIC  (vamp_1.0/vpfits.py:357-358,
// 419-427): PyMC 2 minimises -logp (priors + Normal likelihood) with scipy.optimize.fmin
// (downhill simplex: rho=1, chi=2, psi=0.5, sigma=0.5; stop when max|x_i - x_0| <= xtol and
// max|f_i - f_0| <= ftol, or at maxiter / maxfun) from the last MCMC state, then reports
// BIC = k ln N - 2 lnL with the *likelihood* at the optimum (restated in oracle/sampler.py).
// The objective is evaluated by the same fused model function as K1 (eval_model); the simplex
// lives in shared memory; the whole optimisation runs inside one launch.
#include "vamp_device.cuh"
#include "vamp_kernels.h"

namespace {

constexpr int K3_WARPS = 4;
constexpr double FWHM_PER_SIGMA = 2.35482004503094938202;

__host__ __device__ inline size_t k3_align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T>
__host__ __device__ inline size_t k3_warp_bytes(int kmax, int ts, bool lsf) {
    size_t b = 0;
    b += k3_align16(sizeof(double) * ((size_t)(kmax + 1) * ts + (kmax + 1) + 4 * ts));   // sim, fsim, xbar, xt, xr, xs
    if (lsf) b += k3_align16(sizeof(T) * (VAMP_TILE_PIX + VAMP_MAX_LSF_TAPS));
    return b;
}

template <typename T, int KIND, int PPL>
struct Objective {
    const MapDev& p;
    const RegionSetDev& rs;
    LaneTile<T, PPL> t;
    int P, n, sdi, taps, lane;
    double ubase, lognorm, c_hi, w_hi;
    T* lsf_buf;
    const T* lsf_s;
    double lnL, chi2;      // of the last evaluation

    // -logp of theta (full layout, shared memory); +inf outside the prior support.
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
        T m[PPL];
        eval_model<T, KIND, PPL>(t, th, n, ubase, m, lsf_buf, lsf_s, taps, lane);
        T S, X;
        lane_residuals<T, PPL>(t, m, S, X);
        const double Sd = (double)warp_sum(S), Xd = (double)warp_sum(X);
        chi2 = Xd;
        lnL = p.sd_free ? lnL_sd(Sd, sd, P) : lnL_noise(Xd, lognorm);
        const double f = -(lnL + lp);
        return (f == f) ? f : INFINITY;
    }
};

template <typename T, int KIND, int PPL>
__device__ void nm_run(const MapDev& p, int s, int lane, double* sm, T* lsf_buf, const T* lsf_s) {
    constexpr int NP = (KIND == VAMP_KIND_VOIGT) ? 4 : 3;
    const RegionSetDev& rs = p.rs;
    const int ts = p.tstride;
    Objective<T, KIND, PPL> f{p, rs};
    const int r = p.item_region[s];
    f.P = rs.npix[r];
    f.n = p.item_ncomp[s];
    f.sdi = 4 * p.cmax;
    f.taps = rs.lsf_taps;
    f.lane = lane;
    f.ubase = rs.ubase[rs.tile_off[r]];
    f.lognorm = rs.lognorm[r];
    f.c_hi = (double)(f.P - 1);
    f.w_hi = 0.5 * (double)(f.P - 1) * (KIND == VAMP_KIND_VOIGT ? FWHM_PER_SIGMA : 1.0);
    f.lsf_buf = lsf_buf;
    f.lsf_s = lsf_s;
    load_tile<T, PPL>(f.t, rs, rs.pix_off[r], f.P, lane);

    const int n = f.n;
    const int k = NP * n + (p.sd_free ? 1 : 0);      // free dimension N of scipy
    const int nv = k + 1;
    // free-parameter predicate over the full layout
    auto is_free = [&](int i) -> bool {
        if (i == f.sdi) return p.sd_free != 0;
        return (i < 4 * n) && ((i & 3) < NP);
    };
    // position of the j-th free parameter in the full layout
    auto free_pos = [&](int j) -> int {
        if (j == NP * n) return f.sdi;
        return 4 * (j / NP) + (j % NP);
    };

    double* sim = sm;                       // [nv][ts]
    double* fsim = sim + (size_t)nv * ts;   // [nv]
    double* xbar = fsim + nv;               // [ts]
    double* xt = xbar + ts;                 // [ts] trial point
    double* xr = xt + ts;                   // [ts]
    double* x0 = xr + ts;                   // [ts]

    for (int i = lane; i < ts; i += 32) x0[i] = p.theta[(size_t)s * ts + i];
    __syncwarp();
    // ---- initial simplex (scipy: vertex j+1 perturbs coordinate j by 5 %, or 0.00025 if zero)
    for (int v = 0; v < nv; ++v) {
        const int pj = v > 0 ? free_pos(v - 1) : -1;
        for (int i = lane; i < ts; i += 32) {
            double x = x0[i];
            if (i == pj) x = (x != 0.0) ? 1.05 * x : 0.00025;
            sim[(size_t)v * ts + i] = x;
        }
    }
    __syncwarp();
    int nfev = 0;
    for (int v = 0; v < nv; ++v) {
        const double fv = f(sim + (size_t)v * ts);
        ++nfev;
        if (lane == 0) fsim[v] = fv;
    }
    __syncwarp();

    int iters = 1;
    int ib = 0, iw = 0, iw2 = 0;
    double fb = 0, fw = 0, fw2 = 0;
    const int maxfun = p.maxfun > 0 ? p.maxfun : 0x7fffffff;
    while (true) {
        // ---- order statistics: best, worst, second worst (ties -> lowest index, as a stable sort)
        fb = INFINITY; fw = -INFINITY; fw2 = -INFINITY; ib = iw = iw2 = 0;
        for (int v = 0; v < nv; ++v) {          // tiny (<= 34), warp-uniform
            const double x = fsim[v];
            if (x < fb) { fb = x; ib = v; }
        }
        for (int v = 0; v < nv; ++v) {
            const double x = fsim[v];
            if (x >= fw) { fw = x; iw = v; }     // last among equals = sorted-last
        }
        for (int v = 0; v < nv; ++v) {
            if (v == iw) continue;
            const double x = fsim[v];
            if (x >= fw2) { fw2 = x; iw2 = v; }
        }
        if (!(nfev < maxfun && iters < p.maxiter)) break;
        // ---- convergence test
        double dx = 0.0;
        for (int i = lane; i < ts; i += 32) {
            if (!is_free(i)) continue;
            const double b = sim[(size_t)ib * ts + i];
            for (int v = 0; v < nv; ++v) dx = fmax(dx, fabs(sim[(size_t)v * ts + i] - b));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dx = fmax(dx, __shfl_xor_sync(VAMP_FULL_MASK, dx, o));
        double df = 0.0;
        for (int v = 0; v < nv; ++v) df = fmax(df, fabs(fsim[v] - fb));
        if (dx <= p.xtol && df <= p.ftol) break;

        // ---- centroid of all but the worst
        for (int i = lane; i < ts; i += 32) {
            double a = 0.0;
            for (int v = 0; v < nv; ++v)
                if (v != iw) a += sim[(size_t)v * ts + i];
            xbar[i] = is_free(i) ? a / (double)k : sim[(size_t)ib * ts + i];
        }
        __syncwarp();
        const double* xw = sim + (size_t)iw * ts;
        auto combine = [&](double* dst, double ca, double cb_) {     // dst = ca*xbar + cb*xw
            for (int i = lane; i < ts; i += 32) dst[i] = is_free(i) ? ca * xbar[i] + cb_ * xw[i] : xbar[i];
            __syncwarp();
        };
        auto replace_worst = [&](const double* src, double fv) {
            for (int i = lane; i < ts; i += 32) sim[(size_t)iw * ts + i] = src[i];
            if (lane == 0) fsim[iw] = fv;
            __syncwarp();
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
            combine(xt, 1.5, -0.5);                                 // outside contraction (1 + psi*rho, -psi*rho)
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
                for (int i = lane; i < ts; i += 32) {
                    const double b = sim[(size_t)ib * ts + i];
                    sim[(size_t)v * ts + i] = b + 0.5 * (sim[(size_t)v * ts + i] - b);
                }
                __syncwarp();
                const double fv = f(sim + (size_t)v * ts);
                ++nfev;
                if (lane == 0) fsim[v] = fv;
            }
            __syncwarp();
        }
        ++iters;
    }
    // ---- outputs at the best vertex
    const double fbest = f(sim + (size_t)ib * ts);
    for (int i = lane; i < ts; i += 32) p.theta[(size_t)s * ts + i] = sim[(size_t)ib * ts + i];
    if (lane == 0) {
        p.out_logp[s] = -fbest;
        p.out_lnL[s] = f.lnL;
        p.out_chi2[s] = f.chi2;
        p.out_bic[s] = (double)k * log((double)f.P) - 2.0 * f.lnL;
        p.out_iters[s] = iters;
        p.out_nfev[s] = nfev;
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(K3_WARPS * 32)
k3_map_kernel(MapDev p, int kmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool LSF = p.rs.lsf_taps > 1;
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const size_t head = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS);
    if (LSF) {
        for (int i = threadIdx.x; i < p.rs.lsf_taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(p.rs)[i];
        __syncthreads();
    }
    const int s = blockIdx.x * K3_WARPS + warp;
    if (s >= p.S) return;
    unsigned char* base = smem_raw + head + (size_t)warp * k3_warp_bytes<T>(kmax, p.tstride, LSF);
    double* sm = reinterpret_cast<double*>(base);
    T* lsf_buf = reinterpret_cast<T*>(base + k3_align16(sizeof(double) * ((size_t)(kmax + 1) * p.tstride + (kmax + 1) + 4 * p.tstride)));
    const int P = p.rs.npix[p.item_region[s]];
    const int ppl = (P + 31) >> 5;
    if (ppl <= 1) nm_run<T, KIND, 1>(p, s, lane, sm, lsf_buf, lsf_s);
    else if (ppl <= 2) nm_run<T, KIND, 2>(p, s, lane, sm, lsf_buf, lsf_s);
    else if (ppl <= 4) nm_run<T, KIND, 4>(p, s, lane, sm, lsf_buf, lsf_s);
    else nm_run<T, KIND, 8>(p, s, lane, sm, lsf_buf, lsf_s);
}

}  // namespace

cudaError_t VAMP_INST_NAME(vamp_launch_k3)(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st) {
    typedef InstT T;
    if (p.S <= 0) return cudaSuccess;
    if (max_npix > VAMP_TILE_PIX) return cudaErrorInvalidValue;
    const bool LSF = p.rs.lsf_taps > 1;
    const int NP = (InstKind == VAMP_KIND_VOIGT) ? 4 : 3;
    const int kmax = NP * max_ncomp + 1;
    const size_t smem = k3_align16(sizeof(T) * VAMP_MAX_LSF_TAPS) + K3_WARPS * k3_warp_bytes<T>(kmax, p.tstride, LSF);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = k3_map_kernel<T, InstKind>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(p.S + K3_WARPS - 1) / K3_WARPS, K3_WARPS * 32, smem, st>>>(p, kmax);
    return cudaGetLastError();
}
