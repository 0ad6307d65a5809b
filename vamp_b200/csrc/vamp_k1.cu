// vamp_k1.cu - K1: fused Voigt/Gauss optical depth -> exp(-sum tau) -> LSF -> chi^2 / lnL.
//
// Replaces, per parameter vector ("item"), the reference chain
//   VPfit.VoigtFunction / GaussFunction (vamp_1.0/vpfits.py:44-76) -> total = Tau2flux(sum(...))
//   (vpfits.py:334-338, physics.py:98-105) -> mc.Normal("obs", ...) log-density (vpfits.py:39,341)
//   and VPfit.Chisquared (vpfits.py:109-118).
//
// Two schedules ("ragged batches"):
//   k1_warp : regions of <= 256 pixels.  One warp per item, 8 items per CTA; everything in
//             registers, LSF through a warp-private shared-memory halo buffer.
//   k1_cta  : longer regions.  One CTA per (item, 2048-pixel chunk) with a shared-memory halo
//             of the LSF half-width on both sides; per-chunk partial sums are combined in a
//             fixed order by k1_finalize (bit-reproducible, no atomics).
// Algorithmic HBM bytes: spectrum 12 B/pixel (fp32; 24 B fp64) per item tile + 32 B/component
// + 16 B result per item; flops per pixel-component: see vamp_device.cuh.
#include "vamp_device.cuh"
#include "vamp_kernels.h"

namespace {

constexpr int K1_WARPS = 8;
constexpr int K1_CHUNK = 2048;   // pixels per CTA in k1_cta

template <typename T, int KIND, int PPL, bool MULT>
__device__ __forceinline__ void k1_warp_run(const RegionSetDev& rs, int item, int r, int P, int ncomp,
                                            const double* __restrict__ th, const double* __restrict__ sd,
                                            double* __restrict__ out_lnL, double* __restrict__ out_chi2,
                                            T* __restrict__ out_flux, long long fstride, T* lsf_buf, const T* lsf_s,
                                            int lane) {
    LaneTile<T, PPL> t;
    const int off = rs.pix_off[r];
    load_tile<T, PPL>(t, rs, off, P, lane);
    const double ubase = rs.ubase[rs.tile_off[r]];
    T m[PPL];
    MultD md;
    const bool mult = MULT && region_mult(rs, r, md);
    eval_model<T, KIND, PPL, MULT>(t, th, ncomp, ubase, m, lsf_buf, lsf_s, rs.lsf_taps, lane, mult ? &md : nullptr);
    if (out_flux) {
#pragma unroll
        for (int j = 0; j < PPL; ++j)
            if ((t.valid >> j) & 1u) out_flux[(long long)item * fstride + j * 32 + lane] = m[j];
    }
    T S, X;
    lane_residuals<T, PPL>(t, m, S, X);
    S = warp_sum(S);
    X = warp_sum(X);
    if (lane == 0) {
        if (out_chi2) out_chi2[item] = (double)X;
        if (out_lnL) out_lnL[item] = sd ? lnL_sd((double)S, sd[item], P) : lnL_noise((double)X, rs.lognorm[r]);
    }
}

template <typename T, int KIND, bool MULT>
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_warp_kernel(RegionSetDev rs, int B, const int* __restrict__ item_region, const int* __restrict__ item_ncomp,
               const double* __restrict__ theta, int tstride, const double* __restrict__ sd,
               double* __restrict__ out_lnL, double* __restrict__ out_chi2, T* __restrict__ out_flux,
               long long fstride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* lsf_s = reinterpret_cast<T*>(smem_raw);
    const bool LSF = rs.lsf_taps > 1;
    const int taps = LSF ? rs.lsf_taps : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* lsf_buf = lsf_s + VAMP_MAX_LSF_TAPS + warp * (VAMP_TILE_PIX + VAMP_MAX_LSF_TAPS);
    if (LSF) {
        for (int i = threadIdx.x; i < taps; i += blockDim.x) lsf_s[i] = RS<T>::lsf(rs)[i];
        __syncthreads();
    }
    const int item = blockIdx.x * K1_WARPS + warp;
    if (item >= B) return;
    const int r = item_region[item];
    const int P = rs.npix[r];
    if (P > VAMP_TILE_PIX) return;                       // k1_cta's job
    const int ncomp = item_ncomp[item];
    const double* th = theta + (size_t)item * tstride;
    const int ppl = (P + 31) >> 5;
#define VAMP_K1_RUN(PPL_)                                                                                     \
    k1_warp_run<T, KIND, PPL_, MULT>(rs, item, r, P, ncomp, th, sd, out_lnL, out_chi2, out_flux, fstride, \
                                          lsf_buf, lsf_s, lane)
    if (ppl <= 1) VAMP_K1_RUN(1);
    else if (ppl <= 2) VAMP_K1_RUN(2);
    else if (ppl <= 4) VAMP_K1_RUN(4);
    else VAMP_K1_RUN(8);
#undef VAMP_K1_RUN
}

// ------------------------------------------------------------------------------------------------
template <typename T>
struct CompS {          // component constants in shared memory (centroid kept in double)
    T A, g, y, y2, L2;
};

template <typename T, int KIND, bool MULT>
__global__ void __launch_bounds__(256)
k1_cta_kernel(RegionSetDev rs, int B, int nchunk_max, const int* __restrict__ item_region,
              const int* __restrict__ item_ncomp, const double* __restrict__ theta, int tstride,
              double2* __restrict__ partial, T* __restrict__ out_flux, long long fstride) {
    constexpr int PPL = K1_CHUNK / 256;    // 8
    constexpr int ROWS = PPL + 2;          // + left / right halo tiles
    const int item = blockIdx.x / nchunk_max;
    const int chunk = blockIdx.x - item * nchunk_max;
    if (item >= B) return;
    const int r = item_region[item];
    const int P = rs.npix[r];
    if (P <= VAMP_TILE_PIX) return;        // k1_warp's job
    const int g0 = chunk * K1_CHUNK;
    if (g0 >= P) return;
    const int ncomp0 = item_ncomp[item];
    const double* th = theta + (size_t)item * tstride;
    const bool LSF = rs.lsf_taps > 1;
    const bool FLUX = out_flux != nullptr;
    const int taps = LSF ? rs.lsf_taps : 0;
    const int h = taps >> 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);                  // [16]
    T* lsf_s = reinterpret_cast<T*>(red + 16);                          // [MAX_TAPS]
    T* mbuf = lsf_s + VAMP_MAX_LSF_TAPS;                                // [CHUNK + MAX_TAPS]
    // lines to evaluate: the components, then (multiplet regions) their second lines; from here on `ncomp`
    // counts lines
    MultD md;
    md.fr = md.sh = md.ar = md.lr = 0.0;
    const int nlines = (MULT && region_mult(rs, r, md)) ? 2 * ncomp0 : ncomp0;     // MULT: a separate instantiation
    const int ncomp = nlines;
    CompS<T>* cs = reinterpret_cast<CompS<T>*>(mbuf + K1_CHUNK + VAMP_MAX_LSF_TAPS + 1);  // [nlines]
    T* cb = reinterpret_cast<T*>(cs + nlines);                          // [ROWS][nlines]

    const int off = rs.pix_off[r];
    const int tile0 = rs.tile_off[r] + chunk * PPL;
    const int ntile = (P + VAMP_TILE_PIX - 1) / VAMP_TILE_PIX;
    for (int i = tid; i < taps; i += 256) lsf_s[i] = RS<T>::lsf(rs)[i];
    for (int c = tid; c < nlines; c += 256) {
        double pa = th[VAMP_TH(c, 0)], pce = th[VAMP_TH(c, 1)], pw1 = th[VAMP_TH(c, 2)], pw2 = th[VAMP_TH(c, 3)];
        if (MULT) line_params<KIND>(th, c, ncomp0, md, pa, pce, pw1, pw2);
        const CompC<T> cc = make_comp_k1<T, KIND>(pa, pce, pw1, pw2, 0.0);
        cs[c].A = cc.A; cs[c].g = cc.g; cs[c].y = cc.y; cs[c].y2 = cc.y2; cs[c].L2 = cc.L2;
    }
    for (int i = tid; i < ROWS * nlines; i += 256) {
        const int row = i / nlines, c = i - row * nlines;
        // rows 0..7: tiles of this chunk; row 8: tile left of the chunk; row 9: tile right of it
        int tl = (row < PPL) ? chunk * PPL + row : (row == PPL ? chunk * PPL - 1 : chunk * PPL + PPL);
        tl = min(max(tl, 0), ntile - 1);
        double pa, pce = th[VAMP_TH(c, 1)], pw1, pw2;
        if (MULT) line_params<KIND>(th, c, ncomp0, md, pa, pce, pw1, pw2);
        cb[i] = T(pce - rs.ubase[rs.tile_off[r] + tl]);
    }
    __syncthreads();
    (void)tile0;

    // ---- own pixels: g = g0 + j*256 + tid
    T du[PPL], f[PPL], w[PPL], m[PPL];
    unsigned valid = 0u;
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        const int g = g0 + j * 256 + tid;
        const bool ok = g < P;
        du[j] = ok ? RS<T>::du(rs)[off + g] : T(0);
        f[j] = ok ? RS<T>::f(rs)[off + g] : T(1);
        w[j] = ok ? RS<T>::w(rs)[off + g] : T(0);
        valid |= ok ? (1u << j) : 0u;
    }
    {
        // packed two-row form: rows (2 jj, 2 jj + 1) of this thread travel together, one path per warp and pair
        P2<T> tau2[PPL / 2];
#pragma unroll
        for (int jj = 0; jj < PPL / 2; ++jj) tau2[jj] = p2_bc<T>(T(0));
        unsigned nfar = 0u;
        for (int c = 0; c < ncomp; ++c) {
            CompC<T> cc;
            cc.A = cs[c].A; cc.g = cs[c].g; cc.y = cs[c].y; cc.y2 = cs[c].y2; cc.L2 = cs[c].L2;
            cc.cb = T(0);
            const P2<T> A2 = p2_bc<T>(Real<T>::NEGK() * cc.A);
#pragma unroll
            for (int jj = 0; jj < PPL / 2; ++jj) {
                P2<T> X, S;
                bool far;
                pair_classify<T, KIND>(cc, p2_make(du[2 * jj] - cb[(2 * jj) * ncomp + c], du[2 * jj + 1] - cb[(2 * jj + 1) * ncomp + c]),
                                       X, S, far);
                tau2[jj] = p2_fma(A2, pair_eval<T, KIND>(cc, X, S, far, nfar), tau2[jj]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < PPL / 2; ++jj) {
            const P2<T> M = p2_exp_scaled<T>(tau2[jj]);
            m[2 * jj] = p2_lo(M);
            m[2 * jj + 1] = p2_hi(M);
        }
    }

    if (LSF) {
#pragma unroll
        for (int j = 0; j < PPL; ++j) mbuf[h + j * 256 + tid] = ((valid >> j) & 1u) ? (m[j] - T(1)) : T(0);
        // ---- halo pixels: warps 0..1 evaluate the 2h pixels just outside the chunk
        if (tid < 64) {
            const bool left = tid < h;
            const int i = left ? tid : tid - h;                 // index inside its halo
            const int g = left ? g0 - h + i : g0 + K1_CHUNK + i;
            const bool act = (tid < 2 * h) && g >= 0 && g < P;
            const int row = left ? PPL : PPL + 1;
            T d1[1], H1[1];
            const T duh = act ? RS<T>::du(rs)[off + g] : T(0);
            T tauh = T(0);
            for (int c = 0; c < ncomp; ++c) {
                CompC<T> cc;
                cc.A = cs[c].A; cc.g = cs[c].g; cc.y = cs[c].y; cc.y2 = cs[c].y2; cc.L2 = cs[c].L2;
                cc.cb = T(0);
                d1[0] = duh - cb[row * ncomp + c];
                comp_shape<T, KIND, 1>(cc, d1, H1);
                tauh = Real<T>::fma_(cc.A, H1[0], tauh);
            }
            if (tid < 2 * h) mbuf[left ? i : h + K1_CHUNK + i] = act ? (Real<T>::exp_neg(tauh) - T(1)) : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            T acc = T(1);
            const T* b = mbuf + j * 256 + tid + 2 * h;
            for (int t = 0; t < taps; ++t) acc = Real<T>::fma_(lsf_s[t], b[-t], acc);
            m[j] = acc;
        }
    }
    if (FLUX) {
#pragma unroll
        for (int j = 0; j < PPL; ++j)
            if ((valid >> j) & 1u) out_flux[(long long)item * fstride + g0 + j * 256 + tid] = m[j];
    }
    T S = T(0), X = T(0);
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        const T rr = ((valid >> j) & 1u) ? (f[j] - m[j]) : T(0);
        S = Real<T>::fma_(rr, rr, S);
        const T rw = rr * w[j];
        X = Real<T>::fma_(rw, rw, X);
    }
    double Sd = (double)warp_sum(S), Xd = (double)warp_sum(X);
    if (lane == 0) { red[warp] = Sd; red[8 + warp] = Xd; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += red[k]; b += red[8 + k]; }
        partial[(size_t)item * nchunk_max + chunk] = make_double2(a, b);
    }
}

__global__ void k1_finalize_kernel(RegionSetDev rs, int B, int nchunk_max, const int* __restrict__ item_region,
                                   const double2* __restrict__ partial, const double* __restrict__ sd,
                                   double* __restrict__ out_lnL, double* __restrict__ out_chi2) {
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= B) return;
    const int r = item_region[item];
    const int P = rs.npix[r];
    if (P <= VAMP_TILE_PIX) return;
    const int nchunk = (P + K1_CHUNK - 1) / K1_CHUNK;
    double S = 0.0, X = 0.0;
    for (int k = 0; k < nchunk; ++k) {
        const double2 p = partial[(size_t)item * nchunk_max + k];
        S += p.x;
        X += p.y;
    }
    if (out_chi2) out_chi2[item] = X;
    if (out_lnL) out_lnL[item] = sd ? lnL_sd(S, sd[item], P) : lnL_noise(X, rs.lognorm[r]);
}

template <typename T, int KIND>
cudaError_t launch_k1_t(const RegionSetDev& rs, int B, int max_npix, int max_ncomp, const int* item_region,
                        const int* item_ncomp, const double* theta, int tstride, const double* sd, double* out_lnL,
                        double* out_chi2, void* out_flux, long long fstride, double2* partial, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    const bool LSF = rs.lsf_taps > 1;
    {   // small regions
        const size_t smem = sizeof(T) * (VAMP_MAX_LSF_TAPS + K1_WARPS * (VAMP_TILE_PIX + VAMP_MAX_LSF_TAPS));
        const int grid = (B + K1_WARPS - 1) / K1_WARPS;
        if (rs.mult)      // region set with a multiplet table: the two-lines-per-component instantiation
            k1_warp_kernel<T, KIND, true><<<grid, K1_WARPS * 32, LSF ? smem : 0, st>>>(
                rs, B, item_region, item_ncomp, theta, tstride, sd, out_lnL, out_chi2, (T*)out_flux, fstride);
        else
            k1_warp_kernel<T, KIND, false><<<grid, K1_WARPS * 32, LSF ? smem : 0, st>>>(
                rs, B, item_region, item_ncomp, theta, tstride, sd, out_lnL, out_chi2, (T*)out_flux, fstride);
    }
    if (max_npix > VAMP_TILE_PIX) {
        const int nchunk_max = (max_npix + K1_CHUNK - 1) / K1_CHUNK;
        const size_t nl_max = (size_t)max_ncomp * (rs.mult ? 2 : 1);          // multiplet regions: two lines per component
        const size_t smem = 16 * sizeof(double) + sizeof(T) * (VAMP_MAX_LSF_TAPS + K1_CHUNK + VAMP_MAX_LSF_TAPS + 1) +
                            sizeof(CompS<T>) * nl_max + sizeof(T) * (K1_CHUNK / 256 + 2) * nl_max + 16;
        auto kern = rs.mult ? k1_cta_kernel<T, KIND, true> : k1_cta_kernel<T, KIND, false>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        const long long grid = (long long)B * nchunk_max;
        if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
        kern<<<(unsigned)grid, 256, smem, st>>>(rs, B, nchunk_max, item_region, item_ncomp, theta, tstride, partial,
                                               (T*)out_flux, fstride);
        k1_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(rs, B, nchunk_max, item_region, partial, sd, out_lnL,
                                                             out_chi2);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t VAMP_INST_NAME(vamp_launch_k1)(const RegionSetDev& rs, int B, int max_npix, int max_ncomp,
                                           const int* item_region, const int* item_ncomp, const double* theta,
                                           int tstride, const double* sd, double* out_lnL, double* out_chi2,
                                           void* out_flux, long long fstride, void* partial, cudaStream_t st) {
    return launch_k1_t<InstT, InstKind>(rs, B, max_npix, max_ncomp, item_region, item_ncomp, theta, tstride, sd,
                                        out_lnL, out_chi2, out_flux, fstride, reinterpret_cast<double2*>(partial), st);
}
