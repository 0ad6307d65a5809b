// vamp_cabi.cu - the C ABI of libvamp_b200.so (declared in include/vamp_b200.h).
// Plain CUDA runtime; no torch, no Python.  See the header for the reference interfaces
// (vamp_1.0/vpfits.py file:line) each entry point replaces.
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/vamp_b200.h"
#include "vamp_kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define VAMP_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) return fail(-100, "%s: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct vamp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf partial;     // k1_cta per-chunk partial sums
    DevBuf scratch[8];  // staging for the *_host entry points
    // Region blobs of destroyed region sets, kept for the next upload: cudaFree is a device-wide synchronisation and
    // was measured at 0.4 - 0.8 s on some calls (tools/bench_regions.py: a third of a 512-region fit), so a batched
    // fitter that uploads a region set per batch must not pay it per batch.
    DevBuf region_pool[4];
    void* take_blob(size_t bytes, size_t* cap) {
        int best = -1;
        for (int i = 0; i < 4; ++i)
            if (region_pool[i].p && region_pool[i].cap >= bytes && (best < 0 || region_pool[i].cap < region_pool[best].cap)) best = i;
        if (best < 0) return nullptr;
        void* p = region_pool[best].p;
        *cap = region_pool[best].cap;
        region_pool[best].p = nullptr;
        region_pool[best].cap = 0;
        return p;
    }
    void give_blob(void* p, size_t cap) {
        int slot = -1;
        for (int i = 0; i < 4; ++i)
            if (!region_pool[i].p) { slot = i; break; }
        if (slot < 0) {                       // full: keep the larger blobs
            int smallest = 0;
            for (int i = 1; i < 4; ++i)
                if (region_pool[i].cap < region_pool[smallest].cap) smallest = i;
            if (region_pool[smallest].cap >= cap) { cudaFree(p); return; }
            cudaFree(region_pool[smallest].p);
            slot = smallest;
        }
        region_pool[slot].p = p;
        region_pool[slot].cap = cap;
    }
};

struct vamp_regions {
    vamp_ctx* ctx = nullptr;
    int R = 0;
    int max_npix = 0;
    long long total_pix = 0;   // unpadded
    long long padded_pix = 0;
    std::vector<int> npix, pix_off, tile_off;
    std::vector<double> nu0, scale;
    void* blob = nullptr;      // one device allocation for everything (from / back to the context's pool)
    size_t blob_cap = 0;
    void* mult = nullptr;      // [R][4] multiplet table inside the blob (vamp_regions_set_multiplet)
    RegionSetDev dev{};
};

namespace {

template <typename T>
__global__ void faddeeva_test_kernel(int n, const double* __restrict__ x, const double* __restrict__ y,
                                     double* __restrict__ out) {
    // evaluated through comp_shape so that the branch logic under test is the shipped one
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x);
    const int i = i0 < n ? i0 : n - 1;
    // x = d * g with g = 1, y given directly; L2 chosen so that the very-far branch
    // H = L^2/(L^2 + 4 d^2) equals y^2/(x^2+y^2) (x = 2 sqrt(ln2) d / G, y = sqrt(ln2) L / G).
    CompC<T> cc;
    cc.A = T(1);
    cc.cb = T(0);
    cc.g = T(1);
    const double yc = fmin(y[i], (double)Fad<T>::BIG());      // make_comp clamps the same way
    cc.y = T(yc);
    cc.y2 = cc.y * cc.y;
    cc.L2 = T(4.0 * yc * yc);
    T du[1] = {T(x[i])};
    T H[1];
    comp_shape<T, VAMP_KIND_VOIGT, 1>(cc, du, H);
    if (i0 < n) out[i] = (double)H[0];
}

template <int KIND>
__global__ void profile_kernel(int n, const double* __restrict__ x, double c, double A, double w1, double w2,
                               double* __restrict__ out) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 < n ? i0 : n - 1;
    const CompC<double> cc = make_comp<double, KIND>(A, 0.0, w1, w2, 0.0);
    double du[1] = {x[i] - c};
    double H[1];
    comp_shape<double, KIND, 1>(cc, du, H);
    if (i0 < n) out[i] = A * H[0];
}

__global__ void philox_test_kernel(const unsigned int* c, const unsigned int* k, unsigned int* out) {
    const uint4 r = philox4x32_10(make_uint4(c[0], c[1], c[2], c[3]), make_uint2(k[0], k[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

McmcDev to_dev(const vamp_regions* rg, const vamp_mcmc_args* a) {
    McmcDev p{};
    p.rs = rg->dev;
    p.B = a->n_chains;
    p.item_region = a->item_region;
    p.item_ncomp = a->item_ncomp;
    p.chain_uid = a->chain_uid;
    p.cmax = a->cmax;
    p.tstride = 4 * a->cmax + 1;
    p.sd_free = a->sd_free;
    p.it0 = a->it0;
    p.it1 = a->it1;
    p.burn = a->burn;
    p.thin = a->thin > 0 ? a->thin : 1;
    p.tune_interval = a->tune_interval > 0 ? a->tune_interval : 1000;
    p.tune_throughout = a->tune_throughout;
    p.seed = a->seed;
    p.theta = a->theta;
    p.asf = a->asf;
    p.psd = a->psd;
    p.acc = a->acc;
    p.rej = a->rej;
    p.acc_total = a->acc_total;
    p.logp = a->logp;
    p.lnL = a->lnL;
    p.best_theta = a->best_theta;
    p.best_logp = a->best_logp;
    p.mean = a->mean;
    p.m2 = a->m2;
    p.comom = a->comom;
    p.ntally = a->ntally;
    p.sum_lnL = a->sum_lnL;
    p.sum_lnL2 = a->sum_lnL2;
    p.trace = a->trace;
    p.trace_cap = a->trace_cap;
    p.work = a->work;
    p.am_delay = a->am_delay > 0 ? a->am_delay : 1000;          // PyMC 2 AdaptiveMetropolis defaults
    p.am_interval = a->am_interval > 0 ? a->am_interval : 200;
    p.am = (a->step_kind == VAMP_STEP_ADAPTIVE) ? a->am_state : nullptr;
    return p;
}

// longest region among a launch's chains: the caller's figure (vamp_mcmc_args.max_npix) or the region set's
int launch_npix(const vamp_regions* rg, const vamp_mcmc_args* a) {
    return (a->max_npix > 0 && a->max_npix < rg->max_npix) ? a->max_npix : rg->max_npix;
}

int check_mcmc(const vamp_regions* rg, const vamp_mcmc_args* a) {
    if (!rg || !a) return fail(-1, "null argument");
    if (a->n_chains < 0 || a->cmax < 1 || a->max_ncomp < 1 || a->max_ncomp > a->cmax)
        return fail(-2, "bad sizes: n_chains=%d cmax=%d max_ncomp=%d", a->n_chains, a->cmax, a->max_ncomp);
    const int npix = launch_npix(rg, a);
    if (npix > VAMP_K2_LONG_PIX)
        return fail(-3, "the sampler takes regions of <= %d pixels (got %d); split the region", VAMP_K2_LONG_PIX, npix);
    if (npix > VAMP_TILE_PIX && (rg->dev.mult || a->step_kind == VAMP_STEP_ADAPTIVE))
        return fail(-3, "regions of more than %d pixels (got %d) are sampled by the single-line scalar Metropolis kernels only "
                    "(no multiplet table, no adaptive step method)", VAMP_TILE_PIX, npix);
    if (!a->item_region || !a->item_ncomp || !a->chain_uid || !a->theta || !a->asf || !a->psd || !a->acc || !a->rej ||
        !a->acc_total || !a->logp || !a->lnL || !a->best_theta || !a->best_logp || !a->mean || !a->m2 || !a->comom ||
        !a->ntally || !a->sum_lnL || !a->sum_lnL2)
        return fail(-1, "null state pointer in vamp_mcmc_args");
    if (a->step_kind != VAMP_STEP_METROPOLIS && a->step_kind != VAMP_STEP_ADAPTIVE)
        return fail(-2, "bad step_kind %d", a->step_kind);
    if (a->step_kind == VAMP_STEP_ADAPTIVE) {
        if (!a->am_state) return fail(-1, "step_kind = VAMP_STEP_ADAPTIVE needs am_state [n, cmax, %d]", VAMP_AM_STRIDE);
        if (rg->dev.mult) return fail(-5, "the adaptive step method is not built for multiplet region sets");
    }
    return 0;
}

}  // namespace

namespace {
template <typename T>
__global__ void widen_kernel(const T* in, double* out, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
}  // namespace

extern "C" {

const char* vamp_last_error(void) { return g_err.c_str(); }
int vamp_abi_version(void) { return VAMP_ABI_VERSION; }

int vamp_ctx_create(int device, vamp_ctx** out) {
    if (!out) return fail(-1, "null out");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(-10, "no CUDA device available (%s): vamp_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(-2, "device %d out of range (%d devices)", device, ndev);
    VAMP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VAMP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(-11, "device %d is sm_%d%d; libvamp_b200 is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    vamp_ctx* c = new vamp_ctx();
    c->device = device;
    *out = c;
    return 0;
}

int vamp_ctx_destroy(vamp_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    ctx->partial.release();
    for (auto& b : ctx->scratch) b.release();
    for (auto& b : ctx->region_pool) b.release();
    delete ctx;
    return 0;
}

int vamp_ctx_set_stream(vamp_ctx* ctx, void* s) {
    if (!ctx) return fail(-1, "null ctx");
    ctx->stream = reinterpret_cast<cudaStream_t>(s);
    return 0;
}

int vamp_ctx_synchronize(vamp_ctx* ctx) {
    if (!ctx) return fail(-1, "null ctx");
    VAMP_CUDA(cudaSetDevice(ctx->device));
    VAMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vamp_regions_create(vamp_ctx* ctx, int R, const int* npix, const double* nu, const double* flux,
                        const double* noise, const double* lsf, int lsf_taps, vamp_regions** out) {
    if (!ctx || !out || !npix || !nu || !flux || !noise) return fail(-1, "null argument");
    if (R <= 0) return fail(-2, "R must be positive");
    if (lsf && (lsf_taps < 1 || (lsf_taps & 1) == 0 || lsf_taps > VAMP_MAX_LSF_TAPS))
        return fail(-2, "lsf_taps must be odd and <= %d (got %d)", VAMP_MAX_LSF_TAPS, lsf_taps);
    if (lsf && lsf_taps == 1 && lsf[0] != 1.0) return fail(-2, "a 1-tap LSF must be exactly 1.0 (identity)");
    if (!lsf) lsf_taps = 0;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    vamp_regions* rg = new vamp_regions();
    rg->ctx = ctx;
    rg->R = R;
    rg->npix.assign(npix, npix + R);
    rg->pix_off.resize(R);
    rg->tile_off.resize(R);
    rg->nu0.resize(R);
    rg->scale.resize(R);
    long long poff = 0, toff = 0, src = 0;
    for (int r = 0; r < R; ++r) {
        const int P = npix[r];
        if (P < 1) { delete rg; return fail(-2, "region %d has %d pixels", r, P); }
        rg->pix_off[r] = (int)poff;
        rg->tile_off[r] = (int)toff;
        poff += ((long long)P + 63) / 64 * 64;     // row pairs of the chain kernel never need masking
        toff += ((long long)P + VAMP_TILE_PIX - 1) / VAMP_TILE_PIX;
        rg->max_npix = P > rg->max_npix ? P : rg->max_npix;
        src += P;
        if (poff > 2000000000LL) { delete rg; return fail(-2, "too many pixels in one region set"); }
    }
    rg->total_pix = src;
    rg->padded_pix = poff;
    const long long NP = poff, NT = toff;
    std::vector<float> du32(NP, 0.f), f32(NP, 1.f), w32(NP, 0.f);
    std::vector<double> du64(NP, 0.0), f64(NP, 1.0), w64(NP, 0.0), ubase(NT, 0.0), lognorm(R, 0.0);
    std::vector<float> pdu32(NP), pf32(NP), pw32(NP);      // pair layout (see RegionSetDev)
    std::vector<double> pdu64(NP), pf64(NP), pw64(NP);
    src = 0;
    for (int r = 0; r < R; ++r) {
        const int P = npix[r];
        const double* x = nu + src;
        const double n0 = x[0];
        double s = (P > 1) ? (x[P - 1] - x[0]) / (double)(P - 1) : 1.0;
        if (!(s > 0.0) || !std::isfinite(s)) {
            delete rg;
            return fail(-4, "region %d: frequency grid must be ascending and finite", r);
        }
        rg->nu0[r] = n0;
        rg->scale[r] = s;
        const int ntile = (P + VAMP_TILE_PIX - 1) / VAMP_TILE_PIX;
        double ln = 0.0;
        for (int tl = 0; tl < ntile; ++tl) {
            const int a = tl * VAMP_TILE_PIX;
            const int b = (a + VAMP_TILE_PIX < P) ? a + VAMP_TILE_PIX : P;
            const int mid = (a + b - 1) / 2;
            const double ub = (x[mid] - n0) / s;
            ubase[rg->tile_off[r] + tl] = ub;
            for (int i = a; i < b; ++i) {
                const long long o = rg->pix_off[r] + i;
                const double u = (x[i] - n0) / s;
                du64[o] = u - ub;
                du32[o] = (float)(u - ub);
                f64[o] = flux[src + i];
                f32[o] = (float)flux[src + i];
                // a zero / NaN / inf noise or a non-finite flux would give an inf or NaN weight: chains would freeze
                // (every accept test compares false) and BIC / chi^2 turn NaN without any error raised
                if (!std::isfinite(flux[src + i]) || !std::isfinite(noise[src + i]) || noise[src + i] == 0.0) {
                    delete rg;
                    return fail(-4, "region %d, pixel %d: flux must be finite and noise finite and non-zero (flux %g, noise %g)",
                                r, i, flux[src + i], noise[src + i]);
                }
                const double w = 1.0 / noise[src + i];
                w64[o] = w;
                w32[o] = (float)w;
                ln += std::log(std::fabs(w));
            }
        }
        // padding up to the 64-pixel granule: far away (tau = 0 exactly), flux 1, weight 0, so padded
        // pixels contribute nothing to any sum and kernels may read them unmasked
        for (long long o = rg->pix_off[r] + P; o < rg->pix_off[r] + ((long long)P + 63) / 64 * 64; ++o) {
            du64[o] = 1.0e12;
            du32[o] = 1.0e12f;
        }
        lognorm[r] = ln - 0.5 * (double)P * std::log(2.0 * M_PI);
        src += P;
    }
    for (long long i = 0; i < NP; ++i) {
        // natural pixel i = 64*jj + 32*h + lane  ->  pair element (32*jj + lane), half h
        const long long jj = i / 64, h = (i / 32) & 1, lane = i & 31;
        const long long e = 2 * (32 * jj + lane) + h;
        pdu32[e] = du32[i]; pf32[e] = f32[i]; pw32[e] = w32[i];
        pdu64[e] = du64[i]; pf64[e] = f64[i]; pw64[e] = w64[i];
    }
    std::vector<float> l32(lsf_taps > 0 ? lsf_taps : 1, 1.f);
    std::vector<double> l64(lsf_taps > 0 ? lsf_taps : 1, 1.0);
    for (int i = 0; i < lsf_taps; ++i) { l64[i] = lsf[i]; l32[i] = (float)lsf[i]; }

    // ---- one device blob
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t o_npix = o; o += al(sizeof(int) * R);
    const size_t o_poff = o; o += al(sizeof(int) * R);
    const size_t o_toff = o; o += al(sizeof(int) * R);
    const size_t o_du32 = o; o += al(sizeof(float) * NP);
    const size_t o_f32 = o; o += al(sizeof(float) * NP);
    const size_t o_w32 = o; o += al(sizeof(float) * NP);
    const size_t o_du64 = o; o += al(sizeof(double) * NP);
    const size_t o_f64 = o; o += al(sizeof(double) * NP);
    const size_t o_w64 = o; o += al(sizeof(double) * NP);
    const size_t o_pdu32 = o; o += al(sizeof(float) * NP);
    const size_t o_pf32 = o; o += al(sizeof(float) * NP);
    const size_t o_pw32 = o; o += al(sizeof(float) * NP);
    const size_t o_pdu64 = o; o += al(sizeof(double) * NP);
    const size_t o_pf64 = o; o += al(sizeof(double) * NP);
    const size_t o_pw64 = o; o += al(sizeof(double) * NP);
    const size_t o_ub = o; o += al(sizeof(double) * NT);
    const size_t o_ln = o; o += al(sizeof(double) * R);
    const size_t o_l32 = o; o += al(sizeof(float) * l32.size());
    const size_t o_l64 = o; o += al(sizeof(double) * l64.size());
    const size_t o_mult = o; o += al(sizeof(double) * 4 * (size_t)R);     // room for the multiplet table
    cudaError_t e = cudaSuccess;
    rg->blob = ctx->take_blob(o, &rg->blob_cap);          // a pooled blob of a destroyed region set, if one is large enough
    if (!rg->blob) {
        // a little headroom, so that the next batch of about the same size finds this blob large enough
        const size_t want = o + o / 8;
        e = cudaMalloc(&rg->blob, want);
        if (e != cudaSuccess) { delete rg; return fail(-100, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); }
        rg->blob_cap = want;
    }
    char* base = static_cast<char*>(rg->blob);
    rg->mult = base + o_mult;
    cudaStream_t st = ctx->stream;
#define VAMP_UP(off, vec, bytes)                                                                         \
    do {                                                                                                 \
        e = cudaMemcpyAsync(base + (off), (vec), (bytes), cudaMemcpyHostToDevice, st);                   \
        if (e != cudaSuccess) { ctx->give_blob(rg->blob, rg->blob_cap); delete rg; return fail(-100, "H2D: %s", cudaGetErrorString(e)); } \
    } while (0)
    VAMP_UP(o_npix, rg->npix.data(), sizeof(int) * R);
    VAMP_UP(o_poff, rg->pix_off.data(), sizeof(int) * R);
    VAMP_UP(o_toff, rg->tile_off.data(), sizeof(int) * R);
    VAMP_UP(o_du32, du32.data(), sizeof(float) * NP);
    VAMP_UP(o_f32, f32.data(), sizeof(float) * NP);
    VAMP_UP(o_w32, w32.data(), sizeof(float) * NP);
    VAMP_UP(o_du64, du64.data(), sizeof(double) * NP);
    VAMP_UP(o_f64, f64.data(), sizeof(double) * NP);
    VAMP_UP(o_w64, w64.data(), sizeof(double) * NP);
    VAMP_UP(o_pdu32, pdu32.data(), sizeof(float) * NP);
    VAMP_UP(o_pf32, pf32.data(), sizeof(float) * NP);
    VAMP_UP(o_pw32, pw32.data(), sizeof(float) * NP);
    VAMP_UP(o_pdu64, pdu64.data(), sizeof(double) * NP);
    VAMP_UP(o_pf64, pf64.data(), sizeof(double) * NP);
    VAMP_UP(o_pw64, pw64.data(), sizeof(double) * NP);
    VAMP_UP(o_ub, ubase.data(), sizeof(double) * NT);
    VAMP_UP(o_ln, lognorm.data(), sizeof(double) * R);
    VAMP_UP(o_l32, l32.data(), sizeof(float) * l32.size());
    VAMP_UP(o_l64, l64.data(), sizeof(double) * l64.size());
#undef VAMP_UP
    e = cudaStreamSynchronize(st);   // host vectors go out of scope
    if (e != cudaSuccess) { ctx->give_blob(rg->blob, rg->blob_cap); delete rg; return fail(-100, "sync: %s", cudaGetErrorString(e)); }
    RegionSetDev& d = rg->dev;
    d.R = R;
    d.npix = reinterpret_cast<int*>(base + o_npix);
    d.pix_off = reinterpret_cast<int*>(base + o_poff);
    d.tile_off = reinterpret_cast<int*>(base + o_toff);
    d.du32 = reinterpret_cast<float*>(base + o_du32);
    d.f32 = reinterpret_cast<float*>(base + o_f32);
    d.w32 = reinterpret_cast<float*>(base + o_w32);
    d.du64 = reinterpret_cast<double*>(base + o_du64);
    d.f64 = reinterpret_cast<double*>(base + o_f64);
    d.w64 = reinterpret_cast<double*>(base + o_w64);
    d.pdu32 = reinterpret_cast<float2*>(base + o_pdu32);
    d.pf32 = reinterpret_cast<float2*>(base + o_pf32);
    d.pw32 = reinterpret_cast<float2*>(base + o_pw32);
    d.pdu64 = reinterpret_cast<double2*>(base + o_pdu64);
    d.pf64 = reinterpret_cast<double2*>(base + o_pf64);
    d.pw64 = reinterpret_cast<double2*>(base + o_pw64);
    d.ubase = reinterpret_cast<double*>(base + o_ub);
    d.lognorm = reinterpret_cast<double*>(base + o_ln);
    d.lsf32 = reinterpret_cast<float*>(base + o_l32);
    d.lsf64 = reinterpret_cast<double*>(base + o_l64);
    d.lsf_taps = lsf_taps;
    *out = rg;
    return 0;
}

int vamp_regions_destroy(vamp_regions* rg) {
    if (!rg) return 0;
    cudaSetDevice(rg->ctx->device);
    // kernels of this region set may still be in flight (on the context's stream or on the caller's side streams):
    // the blob goes back to the pool only after they are done
    cudaDeviceSynchronize();
    if (rg->blob) rg->ctx->give_blob(rg->blob, rg->blob_cap);
    delete rg;
    return 0;
}

int vamp_regions_set_multiplet(vamp_regions* rg, const double* freq_ratio, const double* amp_ratio, const double* lorentz_ratio) {
    if (!rg) return fail(-1, "null regions");
    VAMP_CUDA(cudaSetDevice(rg->ctx->device));
    VAMP_CUDA(cudaStreamSynchronize(rg->ctx->stream));       // kernels in flight may still read the old table
    rg->dev.mult = nullptr;
    if (!freq_ratio && !amp_ratio) return 0;                  // cleared: single lines everywhere
    if (!freq_ratio || !amp_ratio) return fail(-1, "freq_ratio and amp_ratio must be given together");
    std::vector<double> t(4 * (size_t)rg->R);
    bool any = false;
    for (int r = 0; r < rg->R; ++r) {
        const double fr = freq_ratio[r], ar = amp_ratio[r], lr = lorentz_ratio ? lorentz_ratio[r] : 1.0;
        if (!(fr > 0.0) || !(ar >= 0.0) || !(lr >= 0.0) || fr != fr || ar != ar || lr != lr || fr > 1e300 || ar > 1e300 || lr > 1e300)
            return fail(-2, "region %d: multiplet ratios must be finite, freq_ratio > 0, amp_ratio >= 0, lorentz_ratio >= 0", r);
        // centroid of line 2 in Hz is fr * c; in the region's scaled units u = (nu - nu0)/s that is fr * u + nu0 (fr - 1)/s
        t[4 * r + 0] = fr;
        t[4 * r + 1] = rg->nu0[r] * (fr - 1.0) / rg->scale[r];
        t[4 * r + 2] = ar;
        t[4 * r + 3] = lr;
        any = any || ar != 0.0;
    }
    if (!any) return 0;
    VAMP_CUDA(cudaMemcpy(rg->mult, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice));   // room inside the blob
    rg->dev.mult = static_cast<const double*>(rg->mult);
    return 0;
}

int vamp_regions_info(const vamp_regions* rg, int* R, int* max_npix, long long* total_pix) {
    if (!rg) return fail(-1, "null regions");
    if (R) *R = rg->R;
    if (max_npix) *max_npix = rg->max_npix;
    if (total_pix) *total_pix = rg->total_pix;
    return 0;
}

int vamp_regions_scale(const vamp_regions* rg, double* nu0, double* scale) {
    if (!rg) return fail(-1, "null regions");
    if (nu0) memcpy(nu0, rg->nu0.data(), sizeof(double) * rg->R);
    if (scale) memcpy(scale, rg->scale.data(), sizeof(double) * rg->R);
    return 0;
}

int vamp_loglike_dev(vamp_ctx* ctx, const vamp_regions* rg, int dtype, int kind, int B, const int* item_region,
                     const int* item_ncomp, int max_ncomp, const double* theta, int theta_stride, const double* sd,
                     double* out_lnL, double* out_chi2, void* out_flux, long long flux_stride) {
    if (!ctx || !rg || !item_region || !item_ncomp || !theta) return fail(-1, "null argument");
    if (dtype != VAMP_F32 && dtype != VAMP_F64) return fail(-2, "bad dtype %d", dtype);
    if (kind != VAMP_GAUSS && kind != VAMP_VOIGT) return fail(-2, "bad kind %d", kind);
    if (B < 0 || max_ncomp < 0 || theta_stride < 4 * max_ncomp) return fail(-2, "bad sizes");
    if (out_flux && flux_stride < rg->max_npix) return fail(-2, "flux_stride %lld < max region size %d", flux_stride, rg->max_npix);
    if (B == 0) return 0;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    void* partial = nullptr;
    if (rg->max_npix > VAMP_TILE_PIX) {
        const int nchunk = (rg->max_npix + VAMP_K1_CHUNK - 1) / VAMP_K1_CHUNK;
        VAMP_CUDA(ctx->partial.reserve(sizeof(double) * 2 * (size_t)B * nchunk));
        partial = ctx->partial.p;
    }
    VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k1, dtype, kind, rg->dev, B, rg->max_npix, max_ncomp, item_region, item_ncomp, theta,
                             theta_stride, sd, out_lnL, out_chi2, out_flux, flux_stride, partial, ctx->stream));
    return 0;
}


int vamp_loglike_host(vamp_ctx* ctx, const vamp_regions* rg, int dtype, int kind, int B, const int* item_region,
                      const int* item_ncomp, int max_ncomp, const double* theta, int theta_stride, const double* sd,
                      double* out_lnL, double* out_chi2, double* out_flux, long long flux_stride) {
    if (!ctx || !rg || !item_region || !item_ncomp || !theta) return fail(-1, "null argument");
    if (B < 0 || max_ncomp < 0 || theta_stride < 4 * max_ncomp) return fail(-2, "bad sizes");
    if (B == 0) return 0;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    // physical (Hz) -> scaled units on the host
    std::vector<double> ths((size_t)B * theta_stride);
    for (int b = 0; b < B; ++b) {
        const int r = item_region[b];
        if (r < 0 || r >= rg->R) return fail(-2, "item %d: region %d out of range", b, r);
        if (item_ncomp[b] < 0 || item_ncomp[b] > max_ncomp) return fail(-2, "item %d: ncomp %d > max_ncomp %d", b, item_ncomp[b], max_ncomp);
        const double n0 = rg->nu0[r], s = rg->scale[r];
        const double* src = theta + (size_t)b * theta_stride;
        double* dst = ths.data() + (size_t)b * theta_stride;
        for (int i = 0; i < theta_stride; ++i) dst[i] = src[i];
        for (int c = 0; c < item_ncomp[b]; ++c) {
            dst[4 * c + 1] = (src[4 * c + 1] - n0) / s;
            dst[4 * c + 2] = src[4 * c + 2] / s;
            dst[4 * c + 3] = src[4 * c + 3] / s;
        }
    }
    cudaStream_t st = ctx->stream;
    const size_t esz = dtype == VAMP_F32 ? sizeof(float) : sizeof(double);
    VAMP_CUDA(ctx->scratch[0].reserve(sizeof(int) * B));
    VAMP_CUDA(ctx->scratch[1].reserve(sizeof(int) * B));
    VAMP_CUDA(ctx->scratch[2].reserve(sizeof(double) * ths.size()));
    VAMP_CUDA(ctx->scratch[3].reserve(sizeof(double) * B));
    VAMP_CUDA(ctx->scratch[4].reserve(sizeof(double) * 2 * B));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, item_region, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, item_ncomp, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[2].p, ths.data(), sizeof(double) * ths.size(), cudaMemcpyHostToDevice, st));
    if (sd) VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[3].p, sd, sizeof(double) * B, cudaMemcpyHostToDevice, st));
    double* d_lnL = static_cast<double*>(ctx->scratch[4].p);
    double* d_chi = d_lnL + B;
    void* d_flux = nullptr;
    if (out_flux) {
        VAMP_CUDA(ctx->scratch[5].reserve(esz * (size_t)B * flux_stride));
        VAMP_CUDA(cudaMemsetAsync(ctx->scratch[5].p, 0, esz * (size_t)B * flux_stride, st));
        d_flux = ctx->scratch[5].p;
    }
    int rc = vamp_loglike_dev(ctx, rg, dtype, kind, B, static_cast<int*>(ctx->scratch[0].p),
                              static_cast<int*>(ctx->scratch[1].p), max_ncomp, static_cast<double*>(ctx->scratch[2].p),
                              theta_stride, sd ? static_cast<double*>(ctx->scratch[3].p) : nullptr, d_lnL, d_chi, d_flux,
                              flux_stride);
    if (rc) return rc;
    if (out_lnL) VAMP_CUDA(cudaMemcpyAsync(out_lnL, d_lnL, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    if (out_chi2) VAMP_CUDA(cudaMemcpyAsync(out_chi2, d_chi, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    if (out_flux) {
        const long long n = (long long)B * flux_stride;
        if (dtype == VAMP_F32) {
            VAMP_CUDA(ctx->scratch[6].reserve(sizeof(double) * n));
            widen_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<float*>(d_flux),
                                                                             static_cast<double*>(ctx->scratch[6].p), n);
            VAMP_CUDA(cudaGetLastError());
            VAMP_CUDA(cudaMemcpyAsync(out_flux, ctx->scratch[6].p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        } else {
            VAMP_CUDA(cudaMemcpyAsync(out_flux, d_flux, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        }
    }
    VAMP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vamp_mcmc_init_dev(vamp_ctx* ctx, const vamp_regions* rg, int kind, const vamp_mcmc_args* a, int from_prior) {
    if (!ctx) return fail(-1, "null ctx");
    if (int rc = check_mcmc(rg, a)) return rc;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    const McmcDev p = to_dev(rg, a);
    if (from_prior) VAMP_CUDA(kind == VAMP_VOIGT ? vamp_launch_init_prior_any_voigt(p, ctx->stream) : vamp_launch_init_prior_any_gauss(p, ctx->stream));
    VAMP_CUDA(vamp_launch_init_state(p, ctx->stream));
    return 0;
}

int vamp_mcmc_run_dev(vamp_ctx* ctx, const vamp_regions* rg, int dtype, int kind, const vamp_mcmc_args* a) {
    if (!ctx) return fail(-1, "null ctx");
    if (int rc = check_mcmc(rg, a)) return rc;
    if (dtype != VAMP_F32 && dtype != VAMP_F64) return fail(-2, "bad dtype %d", dtype);
    if (a->it1 < a->it0 || a->it0 < 0) return fail(-2, "bad iteration range [%d, %d)", a->it0, a->it1);
    VAMP_CUDA(cudaSetDevice(ctx->device));
    const McmcDev p = to_dev(rg, a);
    const int npix = launch_npix(rg, a);
    if (npix > VAMP_TILE_PIX) {
        const cudaError_t e = VAMP_DISPATCH4(vamp_launch_k2l, dtype, kind, p, npix, a->max_ncomp, ctx->stream);
        if (e == cudaErrorInvalidConfiguration)
            return fail(-3, "a %d-pixel region with up to %d components does not fit the sampler's shared memory "
                        "((components + 4) x pixels x %d bytes <= 227 KB): split the region or lower the component count",
                        npix, a->max_ncomp, dtype == VAMP_F64 ? 16 : 8);
        VAMP_CUDA(e);
    } else if (a->step_kind == VAMP_STEP_ADAPTIVE) VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k2a, dtype, kind, p, npix, a->max_ncomp, ctx->stream));
    else if (a->pair_tail && !rg->dev.mult) VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k2t, dtype, kind, p, npix, a->max_ncomp, ctx->stream));
    else if (rg->dev.mult) VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k2m, dtype, kind, p, npix, a->max_ncomp, ctx->stream));
    else VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k2, dtype, kind, p, npix, a->max_ncomp, ctx->stream));
    return 0;
}

int vamp_map_dev(vamp_ctx* ctx, const vamp_regions* rg, int dtype, int kind, const vamp_map_args* a) {
    if (!ctx || !rg || !a) return fail(-1, "null argument");
    if (a->n_starts < 0 || a->cmax < 1 || a->max_ncomp < 1 || a->max_ncomp > a->cmax) return fail(-2, "bad sizes");
    if (!a->item_region || !a->item_ncomp || !a->theta || !a->out_logp || !a->out_lnL || !a->out_chi2 || !a->out_bic ||
        !a->out_iters || !a->out_nfev)
        return fail(-1, "null pointer in vamp_map_args");
    VAMP_CUDA(cudaSetDevice(ctx->device));
    MapDev p{};
    p.rs = rg->dev;
    p.S = a->n_starts;
    p.item_region = a->item_region;
    p.item_ncomp = a->item_ncomp;
    p.cmax = a->cmax;
    p.tstride = 4 * a->cmax + 1;
    p.sd_free = a->sd_free;
    p.maxiter = a->maxiter;
    p.maxfun = a->maxfun;
    p.xtol = a->xtol;
    p.ftol = a->ftol;
    p.theta = a->theta;
    p.out_logp = a->out_logp;
    p.out_lnL = a->out_lnL;
    p.out_chi2 = a->out_chi2;
    p.out_bic = a->out_bic;
    p.out_iters = a->out_iters;
    p.out_nfev = a->out_nfev;
    // (the warp-per-start schedule keeps a region in registers: one coordinate tile; longer regions and LSF region
    //  sets take the CTA schedule, which streams the pixels tile by tile)
    if (dtype == VAMP_MIXED && rg->dev.lsf_taps <= 1 && rg->max_npix <= VAMP_TILE_PIX) {
        VAMP_CUDA(kind == VAMP_VOIGT ? vamp_launch_k3w_mixed_voigt(p, rg->max_npix, a->max_ncomp, ctx->stream)
                                     : vamp_launch_k3w_mixed_gauss(p, rg->max_npix, a->max_ncomp, ctx->stream));
        return 0;
    }
    // (VAMP_MIXED on an LSF or long-region set: the CTA schedule in float64)
    VAMP_CUDA(VAMP_DISPATCH4(vamp_launch_k3, dtype == VAMP_MIXED ? VAMP_F64 : dtype, kind, p, rg->max_npix, a->max_ncomp, ctx->stream));
    return 0;
}

int vamp_detect_dev(vamp_ctx* ctx, int n_spectra, int stride, const int* npix, int max_npix, const double* wavelength,
                    const double* flux, const double* noise, int std_min, int std_max, double* out_det_ratio) {
    if (!ctx || !wavelength || !flux || !noise || !out_det_ratio) return fail(-1, "null argument");
    if (n_spectra < 0 || stride < 1 || max_npix < 1 || max_npix > stride) return fail(-2, "bad sizes: n=%d stride=%d max_npix=%d", n_spectra, stride, max_npix);
    if (std_min < 1 || std_max <= std_min || std_max > 33) return fail(-2, "std range [%d, %d) not in 1 <= std_min < std_max <= 33", std_min, std_max);
    if (n_spectra > 65535) return fail(-2, "at most 65535 sightlines per call (got %d)", n_spectra);
    VAMP_CUDA(cudaSetDevice(ctx->device));
    VAMP_CUDA(vamp_launch_k4(n_spectra, stride, max_npix, npix, wavelength, flux, noise, std_min, std_max, out_det_ratio,
                             ctx->stream));
    return 0;
}

int vamp_detect_host(vamp_ctx* ctx, int n_spectra, int stride, const int* npix, const double* wavelength,
                     const double* flux, const double* noise, int std_min, int std_max, double* out_det_ratio) {
    if (!ctx || !wavelength || !flux || !noise || !out_det_ratio) return fail(-1, "null argument");
    if (n_spectra <= 0) return 0;
    if (stride < 1) return fail(-2, "bad stride %d", stride);
    int max_npix = stride;
    if (npix) {
        max_npix = 0;
        for (int i = 0; i < n_spectra; ++i) {
            if (npix[i] < 0 || npix[i] > stride) return fail(-2, "npix[%d] = %d outside [0, stride]", i, npix[i]);
            max_npix = npix[i] > max_npix ? npix[i] : max_npix;
        }
        if (max_npix == 0) return 0;
    }
    VAMP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nb = sizeof(double) * (size_t)n_spectra * stride;
    for (int k = 0; k < 4; ++k) VAMP_CUDA(ctx->scratch[k].reserve(nb));
    VAMP_CUDA(ctx->scratch[4].reserve(sizeof(int) * (size_t)n_spectra));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, wavelength, nb, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, flux, nb, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[2].p, noise, nb, cudaMemcpyHostToDevice, st));
    if (npix) VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, npix, sizeof(int) * (size_t)n_spectra, cudaMemcpyHostToDevice, st));
    const int rc = vamp_detect_dev(ctx, n_spectra, stride, npix ? (const int*)ctx->scratch[4].p : nullptr, max_npix,
                                   (const double*)ctx->scratch[0].p, (const double*)ctx->scratch[1].p,
                                   (const double*)ctx->scratch[2].p, std_min, std_max, (double*)ctx->scratch[3].p);
    if (rc) return rc;
    VAMP_CUDA(cudaMemcpyAsync(out_det_ratio, ctx->scratch[3].p, nb, cudaMemcpyDeviceToHost, st));
    VAMP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vamp_faddeeva_host(vamp_ctx* ctx, int dtype, int n, const double* x, const double* y, double* out_H) {
    if (!ctx || !x || !y || !out_H) return fail(-1, "null argument");
    if (n <= 0) return 0;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    VAMP_CUDA(ctx->scratch[0].reserve(sizeof(double) * n));
    VAMP_CUDA(ctx->scratch[1].reserve(sizeof(double) * n));
    VAMP_CUDA(ctx->scratch[2].reserve(sizeof(double) * n));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    const int grid = (n + 127) / 128;
    if (dtype == VAMP_F32)
        faddeeva_test_kernel<float><<<grid, 128, 0, st>>>(n, (double*)ctx->scratch[0].p, (double*)ctx->scratch[1].p, (double*)ctx->scratch[2].p);
    else
        faddeeva_test_kernel<double><<<grid, 128, 0, st>>>(n, (double*)ctx->scratch[0].p, (double*)ctx->scratch[1].p, (double*)ctx->scratch[2].p);
    VAMP_CUDA(cudaGetLastError());
    VAMP_CUDA(cudaMemcpyAsync(out_H, ctx->scratch[2].p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    VAMP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vamp_profile_host(vamp_ctx* ctx, int kind, int n, const double* x, double centroid, double amplitude, double w1,
                      double w2, double* out_tau) {
    if (!ctx || !x || !out_tau) return fail(-1, "null argument");
    if (kind != VAMP_GAUSS && kind != VAMP_VOIGT) return fail(-2, "bad kind %d", kind);
    if (n <= 0) return 0;
    VAMP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    VAMP_CUDA(ctx->scratch[0].reserve(sizeof(double) * n));
    VAMP_CUDA(ctx->scratch[1].reserve(sizeof(double) * n));
    VAMP_CUDA(cudaMemcpyAsync(ctx->scratch[0].p, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    const int grid = (n + 127) / 128;
    if (kind == VAMP_VOIGT)
        profile_kernel<VAMP_KIND_VOIGT><<<grid, 128, 0, st>>>(n, (double*)ctx->scratch[0].p, centroid, amplitude, w1, w2, (double*)ctx->scratch[1].p);
    else
        profile_kernel<VAMP_KIND_GAUSS><<<grid, 128, 0, st>>>(n, (double*)ctx->scratch[0].p, centroid, amplitude, w1, w2, (double*)ctx->scratch[1].p);
    VAMP_CUDA(cudaGetLastError());
    VAMP_CUDA(cudaMemcpyAsync(out_tau, ctx->scratch[1].p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    VAMP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vamp_philox_host(vamp_ctx* ctx, const unsigned int* counter, const unsigned int* key, unsigned int* out) {
    if (!ctx || !counter || !key || !out) return fail(-1, "null argument");
    VAMP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    VAMP_CUDA(ctx->scratch[0].reserve(64));
    unsigned int* d = static_cast<unsigned int*>(ctx->scratch[0].p);
    VAMP_CUDA(cudaMemcpyAsync(d, counter, 16, cudaMemcpyHostToDevice, st));
    VAMP_CUDA(cudaMemcpyAsync(d + 4, key, 8, cudaMemcpyHostToDevice, st));
    philox_test_kernel<<<1, 1, 0, st>>>(d, d + 4, d + 8);
    VAMP_CUDA(cudaGetLastError());
    VAMP_CUDA(cudaMemcpyAsync(out, d + 8, 16, cudaMemcpyDeviceToHost, st));
    VAMP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
