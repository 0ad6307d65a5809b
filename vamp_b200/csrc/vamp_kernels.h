// vamp_kernels.h - host-visible launchers of the CUDA kernels (internal; the public
// boundary is include/vamp_b200.h).
#pragma once
#include <cuda_runtime.h>
#include "vamp_device.cuh"

// Every kernel file is compiled once per (precision, profile kind); VAMP_DECL4 declares the four
// suffixed launchers (_f32_gauss, _f32_voigt, _f64_gauss, _f64_voigt) of one kernel.
#define VAMP_DECL4(ret, name, ...)            \
    ret name##_f32_gauss(__VA_ARGS__);        \
    ret name##_f32_voigt(__VA_ARGS__);        \
    ret name##_f64_gauss(__VA_ARGS__);        \
    ret name##_f64_voigt(__VA_ARGS__);
#define VAMP_DISPATCH4(name, dtype, kind, ...)                                                      \
    ((dtype) == 0 ? ((kind) == VAMP_KIND_VOIGT ? name##_f32_voigt(__VA_ARGS__) : name##_f32_gauss(__VA_ARGS__)) \
                  : ((kind) == VAMP_KIND_VOIGT ? name##_f64_voigt(__VA_ARGS__) : name##_f64_gauss(__VA_ARGS__)))

// ---- K1 (vamp_k1.cu)
#define VAMP_K1_CHUNK 2048
VAMP_DECL4(cudaError_t, vamp_launch_k1, const RegionSetDev& rs, int B, int max_npix, int max_ncomp,
           const int* item_region, const int* item_ncomp, const double* theta, int tstride, const double* sd,
           double* out_lnL, double* out_chi2, void* out_flux, long long fstride, void* partial, cudaStream_t st)

// ---- K2 (vamp_k2.cu)
#ifndef VAMP_AM_STRIDE
#define VAMP_AM_STRIDE 40                // doubles of adaptive state per (chain, component): include/vamp_b200.h
#endif
struct McmcDev {
    RegionSetDev rs;
    int B;                               // chains in this launch
    const int* item_region;              // [B]
    const int* item_ncomp;               // [B]
    const unsigned long long* chain_uid; // [B] global chain id (Philox counter words 0,1)
    int cmax;                            // component slots; tstride = 4*cmax + 1
    int tstride;
    int sd_free;                         // 1: sd ~ U(0,1) sampled at theta[4*cmax]; 0: per-pixel noise
    int it0, it1;                        // iterations [it0, it1) run by this launch
    int burn, thin, tune_interval, tune_throughout;
    unsigned long long seed;
    double* theta;                       // [B, tstride] in/out (scaled units)
    float* asf;                          // [B, tstride] adaptive scale factors, in/out
    float* psd;                          // [B, tstride] proposal sd, in
    int* acc;                            // [B, tstride] accepted since last tune, in/out
    int* rej;                            // [B, tstride]
    int* acc_total;                      // [B, tstride] accepted after burn-in (diagnostic), in/out
    double* logp;                        // [B] out: log posterior of the current state
    double* lnL;                         // [B] out: log likelihood of the current state
    double* best_theta;                  // [B, tstride] in/out
    double* best_logp;                   // [B] in/out
    double* mean;                        // [B, tstride] Welford mean of tallied samples, in/out
    double* m2;                          // [B, tstride] Welford sum of squares
    double* comom;                       // [B, cmax, 6] co-moments (A,c)(A,w1)(A,w2)(c,w1)(c,w2)(w1,w2)
    int* ntally;                         // [B] in/out
    double* sum_lnL;                     // [B] in/out: sum of lnL over tallies (for DIC/BPIC)
    double* sum_lnL2;                    // [B]
    double* trace;                       // [B, trace_cap, tstride] or nullptr
    int trace_cap;
    unsigned long long* work;            // [B, 3] or nullptr: shape evals, far-path pixels, likelihood passes
    int am_delay, am_interval;           // adaptive Metropolis (vamp_launch_k2a): PyMC 2 delay / interval
    double* am;                          // [B, cmax, VAMP_AM_STRIDE] adaptive state or nullptr (include/vamp_b200.h)
};
cudaError_t vamp_launch_init_prior_any_gauss(const McmcDev& p, cudaStream_t st);
cudaError_t vamp_launch_init_prior_any_voigt(const McmcDev& p, cudaStream_t st);
cudaError_t vamp_launch_init_state(const McmcDev& p, cudaStream_t st);
VAMP_DECL4(cudaError_t, vamp_launch_k2, const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st)
// pair-tail instantiations (regions whose ceil(P / 64) is odd run a final 64-pixel trip instead of padding to a quad)
VAMP_DECL4(cudaError_t, vamp_launch_k2t, const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st)
// adaptive-Metropolis instantiations (one joint proposal per component from the running covariance)
VAMP_DECL4(cudaError_t, vamp_launch_k2a, const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st)
// multiplet instantiations (region sets with a multiplet table: two tied lines per component)
VAMP_DECL4(cudaError_t, vamp_launch_k2m, const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st)
// long-region instantiations (257 .. VAMP_K2_LONG_PIX pixels: one chain per CTA, all tiles of the region in shared memory)
VAMP_DECL4(cudaError_t, vamp_launch_k2l, const McmcDev& p, int max_npix, int max_ncomp, cudaStream_t st)

// ---- K3 (vamp_k3.cu)
struct MapDev {
    RegionSetDev rs;
    int S;                               // starts
    const int* item_region;              // [S]
    const int* item_ncomp;               // [S]
    int cmax, tstride, sd_free;
    int maxiter, maxfun;                 // scipy.optimize.fmin limits
    double xtol, ftol;
    double* theta;                       // [S, tstride] in: start, out: optimum (scaled units)
    double* out_logp;                    // [S]
    double* out_lnL;                     // [S]
    double* out_chi2;                    // [S] noise-weighted chi^2 at the optimum
    double* out_bic;                     // [S] k ln P - 2 lnL
    int* out_iters;                      // [S]
    int* out_nfev;                       // [S]
};
VAMP_DECL4(cudaError_t, vamp_launch_k3, const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st)
// warp-per-start schedule: fp32 objective, float64 simplex and float64 final evaluation (dtype VAMP_MIXED)
cudaError_t vamp_launch_k3w_mixed_gauss(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st);
cudaError_t vamp_launch_k3w_mixed_voigt(const MapDev& p, int max_npix, int max_ncomp, cudaStream_t st);

// ---- K4 (vamp_k4.cu): detection ratio of sightlines (batched matched filter), float64
cudaError_t vamp_launch_k4(int S, int stride, int max_npix, const int* npix, const double* wavelength, const double* flux,
                           const double* noise, int std_min, int std_max, double* out_det, cudaStream_t st);
