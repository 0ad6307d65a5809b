/* vamp_b200.h - C ABI of libvamp_b200.so: the B200-native (sm_100a) likelihood / MCMC / BIC
 * hot path of sarahappleby/VAMP.
 *
 * The reference (vamp_1.0, pure Python) has no FFI of its own: its hot path is the body of
 * VPfit (vamp_1.0/vpfits.py) delegating to PyMC 2 / astropy / scipy.  Each entry point below
 * names the reference interface it replaces (file:line under /root/reference).  The binding a
 * reference maintainer would add (ctypes) is shown in INTEGRATION.md and shipped as
 * vamp_b200/_lib.py.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; vamp_last_error() gives the message
 *    (thread-local).  Nothing throws, nothing calls back into Python.
 *  - *_host entry points take HOST buffers in the reference's physical units (Hz), copy them
 *    to the device, run on the context's stream, copy results back and synchronise.
 *  - *_dev entry points take caller-owned DEVICE pointers (e.g. torch tensors), are
 *    asynchronous on the context's stream, and use SCALED units: for region r with ascending
 *    frequency grid nu_0..nu_{P-1},  u = (nu - nu_0) / s_r,  s_r = (nu_{P-1} - nu_0)/(P-1),
 *    so centroids live in [0, P-1] and widths are in mean-pixel units (vamp_regions_scale
 *    returns nu_0 and s_r).  Amplitudes and sd are unit-free.
 *  - parameter vectors: theta[item][4*cmax + 1] (double):
 *        theta[4c+0..3] = amplitude, centroid, L_fwhm | sigma, G_fwhm | unused   (c < ncomp)
 *        theta[4*cmax]  = sd (free noise parameter of VPfit(noise=None), vpfits.py:39)
 *    mirroring estimated_variables[c]['amplitude'|'centroid'|'L_fwhm'|'G_fwhm'|'sigma']
 *    (vpfits.py:232-262, 276-307).
 */
#ifndef VAMP_B200_H
#define VAMP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VAMP_ABI_VERSION 5   /* 2: vamp_detect_*, VAMP_MIXED for vamp_map_dev; 3: vamp_regions_set_multiplet;
                                4: vamp_mcmc_args.step_kind / am_* (adaptive Metropolis) and pair_tail;
                                5: regions of up to VAMP_MAX_FIT_PIXELS pixels in vamp_mcmc_* / vamp_map_dev,
                                   vamp_mcmc_args.max_npix */
#define VAMP_MAX_FIT_PIXELS 2048   /* longest region the sampler takes (the reference fits arrays of any length, e.g.
                                      2000 pixels in vamp_1.0/vpfits_intro.ipynb); vamp_map_dev (CTA schedule) and
                                      vamp_loglike_* take any length */

enum { VAMP_F32 = 0, VAMP_F64 = 1,            /* arithmetic of the pixel math */
       VAMP_MIXED = 2 };                       /* vamp_map_dev only: fp32 objective during the simplex search,
                                                  float64 simplex and float64 lnL / chi^2 / BIC at the optimum,
                                                  one warp per start (the schedule for thousands of starts) */
enum { VAMP_GAUSS = 0, VAMP_VOIGT = 1 };      /* VPfit(voigt=False|True) */

typedef struct vamp_ctx vamp_ctx;
typedef struct vamp_regions vamp_regions;

const char* vamp_last_error(void);
int vamp_abi_version(void);

/* One context per process/GPU (holds device id, stream, scratch). */
int vamp_ctx_create(int device, vamp_ctx** out);
int vamp_ctx_destroy(vamp_ctx* ctx);
int vamp_ctx_set_stream(vamp_ctx* ctx, void* cuda_stream);   /* cudaStream_t; NULL = default stream */
int vamp_ctx_synchronize(vamp_ctx* ctx);

/* Upload R absorption regions (what VPregion.__init__ holds: frequency_array, flux_array,
 * noise_array - vamp_1.0/vpregion.py:10-19; built by vpspectrum.py:273-279).  HOST arrays,
 * packed back to back, nu ascending inside each region.  lsf: odd number of taps or NULL
 * (identity = the reference).  Precomputes, in float64 on the host, the scaled pixel
 * coordinates relative to a per-256-pixel tile base for both precisions. */
int vamp_regions_create(vamp_ctx* ctx, int R, const int* npix, const double* nu, const double* flux,
                        const double* noise, const double* lsf, int lsf_taps, vamp_regions** out);
int vamp_regions_destroy(vamp_regions* regions);
int vamp_regions_info(const vamp_regions* regions, int* R, int* max_npix, long long* total_pix);
int vamp_regions_scale(const vamp_regions* regions, double* nu0 /*[R]*/, double* scale /*[R]*/);

/* Multiplets (doublets with tied v / b: BASELINE config C3 - CIV 1548/1550, MgII 2796/2803; the reference has
 * only single lines, VPfit.VoigtFunction vpfits.py:58-76 once per component).  In region r every component
 * (amplitude A, centroid c, L_fwhm, G_fwhm | sigma) carries a second line with
 *     amplitude amp_ratio[r] A,  centroid freq_ratio[r] c,  G_fwhm freq_ratio[r] G_fwhm  (sigma likewise),
 *     L_fwhm lorentz_ratio[r] L_fwhm
 * i.e. for lines of rest frequencies nu_1, nu_2 at one velocity and one Doppler width: freq_ratio = nu_2 / nu_1,
 * amp_ratio = (f_2 lambda_2)/(f_1 lambda_1), lorentz_ratio = Gamma_2 / Gamma_1 (NULL = 1).  The free parameters,
 * their priors, BIC's k and every buffer layout are unchanged; K1 / K2 / K3 evaluate both lines.
 * amp_ratio[r] = 0: region r keeps single lines.  HOST arrays [R]; NULL, NULL, NULL clears the table.
 * Synchronises the context's stream. */
int vamp_regions_set_multiplet(vamp_regions* regions, const double* freq_ratio, const double* amp_ratio,
                               const double* lorentz_ratio);

/* K1 - fused model + likelihood.  Replaces, per item:
 *   VPfit.VoigtFunction / GaussFunction (vpfits.py:44-76), total = Tau2flux(sum(profiles))
 *   (vpfits.py:334-338, physics.py:98-105), the Normal("obs") log-density (vpfits.py:39,341;
 *   per-pixel-noise form vamp_2.0/vamp_src/fit/fit.py:127-171) and VPfit.Chisquared
 *   (vpfits.py:109-118).
 * sd != NULL: lnL with the scalar noise parameter sd[item]; sd == NULL: lnL with the region's
 * noise array.  out_chi2 always uses the noise array.  Any out_* may be NULL.
 * out_flux[item*flux_stride + pixel] receives the model flux (the `total` deterministic). */
int vamp_loglike_dev(vamp_ctx* ctx, const vamp_regions* regions, int dtype, int kind, int n_items,
                     const int* item_region, const int* item_ncomp, int max_ncomp, const double* theta,
                     int theta_stride, const double* sd, double* out_lnL, double* out_chi2,
                     void* out_flux /* float* or double* per dtype */, long long flux_stride);
/* Same, HOST buffers, theta in Hz (physical units); out_flux is double regardless of dtype. */
int vamp_loglike_host(vamp_ctx* ctx, const vamp_regions* regions, int dtype, int kind, int n_items,
                      const int* item_region, const int* item_ncomp, int max_ncomp, const double* theta,
                      int theta_stride, const double* sd, double* out_lnL, double* out_chi2,
                      double* out_flux, long long flux_stride);

/* K2 - batched Metropolis chains, one warp per chain.  Replaces mcmc.sample(iter, burn, thin)
 * (vpfits.py:393, 423, 425) i.e. PyMC 2's scalar Metropolis steppers + tune().  All pointers
 * are DEVICE pointers owned by the caller; state persists across calls so a chain can be
 * advanced in slices [it0, it1). */
typedef struct vamp_mcmc_args {
    int n_chains;
    int cmax;                   /* component slots; theta_stride must be 4*cmax + 1 */
    int max_ncomp;              /* max over item_ncomp (sizes shared memory) */
    int sd_free;                /* 1 = VPfit(noise=None): sd ~ U(0,1) sampled; 0 = noise array */
    int it0, it1;               /* iterations to run: [it0, it1) */
    int burn, thin;             /* tally when i >= burn and i % thin == 0 (PyMC 2) */
    int tune_interval;          /* PyMC 2 default 1000 */
    int tune_throughout;        /* PyMC 2 default 1; 0 = adapt during burn-in only */
    unsigned long long seed;    /* Philox key */
    const int* item_region;     /* [n] */
    const int* item_ncomp;      /* [n] */
    const unsigned long long* chain_uid; /* [n] global chain ids (Philox counter) */
    double* theta;              /* [n, 4*cmax+1] in/out, scaled units */
    float* asf;                 /* [n, stride] adaptive scale factors */
    float* psd;                 /* [n, stride] proposal sd (PyMC 2: |initial value|, 1 if 0) */
    int* acc;                   /* [n, stride] accepted since last tune */
    int* rej;                   /* [n, stride] */
    int* acc_total;             /* [n, stride] accepted after burn-in */
    double* logp;               /* [n] log posterior of current state (without uniform constants) */
    double* lnL;                /* [n] */
    double* best_theta;         /* [n, stride] highest-logp state seen at an iteration end */
    double* best_logp;          /* [n] */
    double* mean;               /* [n, stride] Welford mean of tallies   (mcmc.stats(), vpspectrum.py:400) */
    double* m2;                 /* [n, stride] Welford M2 */
    double* comom;              /* [n, cmax, 6] co-moments (A,c)(A,w1)(A,w2)(c,w1)(c,w2)(w1,w2) (chain_covariance, vpfits.py:432-456) */
    int* ntally;                /* [n] */
    double* sum_lnL;            /* [n] sum of lnL over tallies (DIC/BPIC) */
    double* sum_lnL2;           /* [n] */
    double* trace;              /* [n, trace_cap, stride] or NULL  (mcmc.trace(name)[:]) */
    int trace_cap;
    unsigned long long* work;   /* [n, 3] or NULL, accumulated: component evaluations (each over the region's
                                   pixels), far-path pixels among them, likelihood passes - the
                                   algorithmic work bench.py's roofline is computed from */
    /* step method (ABI 4).  VAMP_STEP_METROPOLIS: PyMC 2's scalar Metropolis per stochastic with tune() - what the
     * reference actually runs (mcmc_fit names mc.AdaptiveMetropolis but never installs it, vpfits.py:361, 382-389).
     * VAMP_STEP_ADAPTIVE: PyMC 2's AdaptiveMetropolis (Haario et al. 2001) as `mcmc.use_step_method(
     * mc.AdaptiveMetropolis, [the stochastics of one component])` would install it: one joint Normal proposal per
     * component from the running covariance of its chain history (delay, interval, greedy tallying, scaling
     * 2.4^2 / d, epsilon 1e-5, C_0 = diag(|value| / 100)); sd keeps its scalar Metropolis step. */
    int step_kind;
    int am_delay;               /* PyMC 2 default 1000: first covariance update after this many iterations */
    int am_interval;            /* PyMC 2 default 200: iterations between covariance updates */
    int pair_tail;              /* 1: some chain of this launch fits a region with an odd number of 64-pixel row pairs
                                   (<= 64 or 129..192 pixels): run the pair-tail kernel, which sweeps ceil(P / 64) pairs
                                   instead of whole 128-pixel quads.  Chains with an even count compute the same
                                   either way; a caller must set it whenever such a chain is present (always, for
                                   results that do not depend on batching).  Ignored for multiplet / adaptive runs. */
    int max_npix;               /* ABI 5: longest region among THIS launch's chains, 0 = the region set's longest.  Chains of
                                   regions above 256 pixels run the long-region kernels (one chain per CTA, all
                                   coordinate tiles of the region in shared memory: (max_ncomp + 4) x pixels x 8 bytes
                                   (float64: 16) must fit 227 KB, error -3 otherwise; single lines and the scalar step
                                   method only); a caller that launches its short regions separately keeps them on
                                   the one-tile kernels. */
    double* am_state;           /* [n, cmax, VAMP_AM_STRIDE] adaptive state (NULL unless step_kind = VAMP_STEP_ADAPTIVE):
                                   chain mean [4] | covariance, packed lower triangle [10] | its Cholesky factor [10] |
                                   sum x [4] | sum x x^T [10] | samples in the covariance | samples since the update */
} vamp_mcmc_args;
#define VAMP_STEP_METROPOLIS 0
#define VAMP_STEP_ADAPTIVE 1
#define VAMP_AM_STRIDE 40

/* from_prior != 0: draw PyMC 2-style initial values (xexp = 0.5, Uniforms from their priors,
 * vpfits.py:239-252, 283-297) into theta; then (always) reset asf/psd/counters/statistics. */
int vamp_mcmc_init_dev(vamp_ctx* ctx, const vamp_regions* regions, int kind, const vamp_mcmc_args* a, int from_prior);
int vamp_mcmc_run_dev(vamp_ctx* ctx, const vamp_regions* regions, int dtype, int kind, const vamp_mcmc_args* a);

/* K3 - batched Nelder-Mead MAP + BIC (one CTA per start; one warp per start with VAMP_MIXED).  Replaces
 * mc.MAP(model).fit(iterlim, tol) and map.BIC (vpfits.py:357-358, 419-427). */
typedef struct vamp_map_args {
    int n_starts;
    int cmax, max_ncomp, sd_free;
    int maxiter, maxfun;        /* scipy.optimize.fmin limits; maxfun <= 0: unlimited */
    double xtol, ftol;          /* PyMC 2 MAP.fit: ftol = tol (1e-3), xtol = 1e-4 */
    const int* item_region;     /* [n] */
    const int* item_ncomp;      /* [n] */
    double* theta;              /* [n, 4*cmax+1] in: start, out: optimum (scaled units) */
    double* out_logp;           /* [n] */
    double* out_lnL;            /* [n] */
    double* out_chi2;           /* [n] noise-weighted chi^2 (VPfit.Chisquared) at the optimum */
    double* out_bic;            /* [n] k ln P - 2 lnL, k = (3|4) ncomp + sd_free */
    int* out_iters;             /* [n] */
    int* out_nfev;              /* [n] */
} vamp_map_args;
int vamp_map_dev(vamp_ctx* ctx, const vamp_regions* regions, int dtype, int kind, const vamp_map_args* a);

/* K4 - detection ratio of whole sightlines (batched matched filter).  Replaces the equivalent-width
 * loops and the 2 x 9 Gaussian convolutions of VPspectrum.compute_detection_regions
 * (vamp_1.0/vpspectrum.py:92-122; the vectorised twin is vamp_2.0/vamp_src/dataset/preprocess.py:45-126):
 * out[s, i] = max over std in [std_min, std_max) of flux_func / sqrt(noise_func), -inf at the two end
 * pixels.  float64.  [n_spectra, stride] row-major arrays; npix[s] <= stride pixels of row s are
 * used (NULL: all rows have `stride` pixels).  The sequential region selection
 * (vpspectrum.py:124-174) is host logic (vamp_b200/vpspectrum.py:select_regions). */
int vamp_detect_dev(vamp_ctx* ctx, int n_spectra, int stride, const int* npix, int max_npix, const double* wavelength,
                    const double* flux, const double* noise, int std_min, int std_max, double* out_det_ratio);
int vamp_detect_host(vamp_ctx* ctx, int n_spectra, int stride, const int* npix, const double* wavelength,
                     const double* flux, const double* noise, int std_min, int std_max, double* out_det_ratio);

/* One component's optical depth tau(x) on an arbitrary HOST grid, evaluated in float64 on the
 * device: VPfit.VoigtFunction(x, centroid, amplitude, L_fwhm, G_fwhm) (vpfits.py:58-76; kind =
 * VAMP_VOIGT, w1 = L_fwhm, w2 = G_fwhm) or VPfit.GaussFunction(x, amplitude, centroid, sigma)
 * (vpfits.py:44-54; kind = VAMP_GAUSS, w1 = sigma).  Physical units. */
int vamp_profile_host(vamp_ctx* ctx, int kind, int n, const double* x, double centroid, double amplitude,
                      double w1, double w2, double* out_tau);

/* Device Faddeeva test hook: H(x, y) = sqrt(pi) y Re w(x + iy) for n HOST points (both
 * precisions).  Used by the parity tests against scipy.special.wofz / mpmath. */
int vamp_faddeeva_host(vamp_ctx* ctx, int dtype, int n, const double* x, const double* y, double* out_H);
/* Philox4x32-10 known-answer hook (HOST buffers): out[4] for counter[4], key[2]. */
int vamp_philox_host(vamp_ctx* ctx, const unsigned int* counter, const unsigned int* key, unsigned int* out);

#ifdef __cplusplus
}
#endif
#endif /* VAMP_B200_H */
