// pyramid.cu -- complex steerable pyramid (decompose / reconstruct) for sm_100a.
//
// Replaces steerable.SCFpyr_PyTorch.build / .reconstruct (third party, absent from the reference;
// call sites src/train/pyramid.py:28-33,37,44) fused with Pyramid.coeff_to_values /
// values_to_coeff (src/train/pyramid.py:48-112).  Algorithm: SURVEY.md Appendix A (restated in
// oracle/steerable_shim.py).  B200-first formulation:
//
//  * FLATTENED recursion.  The reference recursion (mask, band IFFT, crop, mask, ...) is linear in
//    the image spectrum X, so band(l,b) = IFFT2_{h_l x w_l}( X[k] * D_l[k] * A_b[k] * (-i)^(nb-1) )
//    with one radial table D_l = lo0 * prod_{j<l} lomask_j * himask_l per level (plan, built on the
//    host in double with the same LUT interpolation the oracle uses) and the angular factor
//    A_b = 2 sqrt(c) cos(theta - pi b/nb)^(nb-1) [cos > 0] evaluated in closed form from the
//    frequency coordinates (no atan2, no table).  Every level depends only on X, so all levels are
//    independent launches; reconstruction sums the level spectra in one gather.
//  * Every 1-D transform (any length: 764 = 4*191, 1358 = 2*7*97, 241, ...) is a shared-memory
//    Stockham FFT (fft_smem.cuh); rows are transformed in batches of rows, columns in tiles of
//    CT adjacent columns, so global traffic is one read + one write of the array per pass.
//  * Fused epilogues/prologues: amplitude |z|, phase atan2(im,re), the per-(level,plane) amplitude
//    maximum (PhaseNet.normalize_vals, src/phase_net/phase_net.py:47-59) are produced by the last
//    row pass of the band IFFT -- the complex band never goes to HBM; reconstruction reads
//    (phase, amplitude) and forms A*(cos,sin) in the first row pass (pyramid.py:103-108).
#include <math.h>

#include <algorithm>
#include <map>
#include <vector>

#include "common.cuh"
#include "fft_smem.cuh"

namespace fvfi {

constexpr int MAX_LEVELS = 40;
constexpr int MAX_BANDS = 8;
constexpr int ROW_ELEMS = 4096;   // complex elements per CTA in a row pass  (2 buffers = 64 KB)
constexpr int COL_ELEMS = 8704;   // complex elements per CTA in a column pass (3 buffers = 204 KB)

struct LevelGeom {
    int h, w;
    size_t off;          // offset (complex elements, per plane) of this level's spectrum in region C
    float rad2_lo, rad2_hi;  // radial support of D_l in normalised radius^2 (for the gather)
};

struct AngParams {
    int nb, order;
    float cs[MAX_BANDS], sn[MAX_BANDS];
    float scale;      // 2*sqrt(const) (build, one-sided) or sqrt(const) (reconstruct, two-sided)
    float2 fac;       // (-i)^(nb-1) (build) or (i)^(nb-1) (reconstruct)
    int one_sided;
    float inv_hh, inv_hw;  // 2/H, 2/W of the FULL image (grid coordinates, prepare_grid)
};

}  // namespace fvfi

struct fvfi_pyr_plan {
    int H, W, height, nbands, L;
    double scale;
    std::vector<fvfi::LevelGeom> lv;            // L band levels + low residual (index L)
    std::vector<fvfi::Fft1D> fy, fx;            // per level (index L = low)
    std::vector<const float*> radial;           // device, per level (index L = low-pass product)
    const float* hi0 = nullptr;                 // device [H*W], unshifted
    std::map<int, float2*> tw;                  // device twiddle tables by length
    std::vector<void*> owned;
    size_t level_elems = 0;                     // sum_l h_l*w_l  (l = 0..L)
    fvfi::AngParams ang_build, ang_rec;
};

namespace fvfi {

// ------------------------------------------------------------------------------------------------
// host: plan
// ------------------------------------------------------------------------------------------------
static int next_size(int n, double s) { return (int)ceil((n - 0.5) / s - 1e-9); }

static void factorize(int n, Fft1D& f) {
    std::vector<int> primes;
    int rem = n;
    for (int d = 2; d * d <= rem; ++d)
        while (rem % d == 0) { primes.push_back(d); rem /= d; }
    if (rem > 1) primes.push_back(rem);
    int twos = 0;
    std::vector<int> others;
    for (int p : primes) { if (p == 2) ++twos; else others.push_back(p); }
    std::sort(others.begin(), others.end(), [](int a, int b) { return a > b; });
    f.n = n;
    f.nfac = 0;
    for (int p : others) f.fac[f.nfac++] = p;
    for (int i = 0; i < twos / 2; ++i) f.fac[f.nfac++] = 4;
    if (twos & 1) f.fac[f.nfac++] = 2;
}

// numpy.interp on an increasing abscissa (end-clamped) -- what upstream's pointOp does.
static double interp(double x, const std::vector<double>& X, const std::vector<double>& Y) {
    const int n = (int)X.size();
    if (x <= X[0]) return Y[0];
    if (x >= X[n - 1]) return Y[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (X[mid] <= x) lo = mid; else hi = mid;
    }
    const double slope = (Y[hi] - Y[lo]) / (X[hi] - X[lo]);
    return slope * (x - X[lo]) + Y[lo];
}

static inline int sfreq_h(int k, int n) { return k < (n + 1) / 2 ? k : k - n; }

template <typename T>
static int upload(fvfi_pyr_plan* p, const std::vector<T>& v, const T** out) {
    void* d = nullptr;
    FVFI_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
    p->owned.push_back(d);
    FVFI_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return FVFI_OK;
}

static int make_fft(fvfi_pyr_plan* p, int n, Fft1D& f) {
    factorize(n, f);
    if (f.nfac > FFT_MAX_FACTORS) { set_error("pyramid: too many FFT factors for n=%d", n); return FVFI_EINVAL; }
    auto it = p->tw.find(n);
    if (it == p->tw.end()) {
        std::vector<float2> t(n);
        for (int k = 0; k < n; ++k) {
            const double a = -2.0 * M_PI * (double)k / (double)n;
            t[k] = make_float2((float)cos(a), (float)sin(a));
        }
        const float2* d = nullptr;
        if (int rc = upload(p, t, &d)) return rc;
        p->tw[n] = (float2*)d;
        it = p->tw.find(n);
    }
    f.tw = it->second;
    return FVFI_OK;
}

static int build_plan(fvfi_pyr_plan* p) {
    const int H = p->H, W = p->W, L = p->L, nb = p->nbands;
    const double s = p->scale, dlt = log2(s);
    // level sizes (SURVEY.md Appendix A.4 / oracle.steerable_shim.level_sizes)
    p->lv.resize(L + 1);
    p->fy.resize(L + 1);
    p->fx.resize(L + 1);
    p->radial.resize(L + 1);
    int h = H, w = W;
    size_t off = 0;
    for (int l = 0; l <= L; ++l) {
        p->lv[l].h = h;
        p->lv[l].w = w;
        p->lv[l].off = off;
        off += (size_t)h * w;
        if (int rc = make_fft(p, h, p->fy[l])) return rc;
        if (int rc = make_fft(p, w, p->fx[l])) return rc;
        h = next_size(h, s);
        w = next_size(w, s);
    }
    p->level_elems = off;
    if (p->lv[L].h < 2 || p->lv[L].w < 2) { set_error("pyramid: height %d too large for %dx%d", p->height, H, W); return FVFI_EINVAL; }

    // raised-cosine tables (upstream rcosFn(1, -0.5))
    const int NT = 259;
    std::vector<double> Xr(NT), Yr(NT), YIr(NT);
    for (int i = 0; i < NT; ++i) {
        const double X0 = M_PI * (double)(i - 257) / 512.0;
        double Y = cos(X0) * cos(X0);
        Yr[i] = Y;
        Xr[i] = -0.5 + 2.0 / M_PI * (X0 + M_PI / 4.0);
    }
    Yr[0] = Yr[1];
    Yr[NT - 1] = Yr[NT - 2];
    for (int i = 0; i < NT; ++i) {
        Yr[i] = sqrt(Yr[i]);
        YIr[i] = sqrt(fabs(1.0 - Yr[i] * Yr[i]));
    }
    auto shifted = [&](double d) { std::vector<double> X(Xr); for (auto& x : X) x -= d; return X; };
    std::vector<std::vector<double>> Xlev(L + 1);
    for (int j = 0; j <= L; ++j) Xlev[j] = shifted((j + 1) * dlt);

    auto log_rad = [&](int fy, int fx) {
        double xv = fx * 2.0 / W, yv = fy * 2.0 / H;
        if (fy == 0 && fx == 0) xv = -2.0 / W;  // prepare_grid: DC sample replaced by its left neighbour
        return log2(sqrt(xv * xv + yv * yv));
    };
    auto clean = [](double v) { return fabs(v) < 1e-12 ? 0.0 : v; };

    // hi0 on the full grid
    {
        std::vector<float> t((size_t)H * W);
        for (int ky = 0; ky < H; ++ky)
            for (int kx = 0; kx < W; ++kx)
                t[(size_t)ky * W + kx] = (float)clean(interp(log_rad(sfreq_h(ky, H), sfreq_h(kx, W)), Xr, Yr));
        if (int rc = upload(p, t, &p->hi0)) return rc;
    }
    // D_l = lo0 * prod_{j<l} lomask_j * himask_l ;  low = lo0 * prod_{j<L} lomask_j
    for (int l = 0; l <= L; ++l) {
        const int hl = p->lv[l].h, wl = p->lv[l].w;
        std::vector<float> t((size_t)hl * wl);
        for (int ky = 0; ky < hl; ++ky)
            for (int kx = 0; kx < wl; ++kx) {
                const double lr = log_rad(sfreq_h(ky, hl), sfreq_h(kx, wl));
                double v = interp(lr, Xr, YIr);
                for (int j = 0; j < l && v != 0.0; ++j) v *= interp(lr, Xlev[j], YIr);
                if (l < L) v *= interp(lr, Xlev[l], Yr);
                t[(size_t)ky * wl + kx] = (float)clean(v);
            }
        if (int rc = upload(p, t, &p->radial[l])) return rc;
        // radial support (log2 radius): himask_l > 0 above -(l+1)dlt - 1 ; lo-product > 0 below -l*dlt (0 for l=0)
        const double lo = (l < L) ? -(l + 1) * dlt - 1.0 : -1e30;
        const double hi = -(double)l * dlt;
        p->lv[l].rad2_lo = (l < L) ? (float)(pow(2.0, 2.0 * lo) * (1.0 - 1e-4)) : -1.f;
        p->lv[l].rad2_hi = (float)(pow(2.0, 2.0 * hi) * (1.0 + 1e-4));
    }
    // angular parameters
    const int order = nb - 1;
    double fo = 1, f2o = 1;
    for (int i = 2; i <= order; ++i) fo *= i;
    for (int i = 2; i <= 2 * order; ++i) f2o *= i;
    const double cst = pow(2.0, 2.0 * order) * fo * fo / (nb * f2o);
    AngParams a{};
    a.nb = nb;
    a.order = order;
    for (int b = 0; b < nb; ++b) { a.cs[b] = (float)cos(M_PI * b / nb); a.sn[b] = (float)sin(M_PI * b / nb); }
    a.inv_hh = (float)(2.0 / H);
    a.inv_hw = (float)(2.0 / W);
    // (-i)^(nb-1) and (i)^(nb-1)
    const float2 pw_m[4] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}};
    const float2 pw_p[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
    p->ang_build = a;
    p->ang_build.scale = (float)(2.0 * sqrt(cst));
    p->ang_build.fac = pw_m[order & 3];
    p->ang_build.one_sided = 1;
    p->ang_rec = a;
    p->ang_rec.scale = (float)sqrt(cst);
    p->ang_rec.fac = pw_p[order & 3];
    p->ang_rec.one_sided = 0;
    return FVFI_OK;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sfreq(int k, int n) { return k < ((n + 1) >> 1) ? k : k - n; }
__device__ __forceinline__ int wrapi(int f, int n) { return f < 0 ? f + n : f; }

// angular factor of band b at signed frequency (fy, fx) of the full grid, times the band's complex constant
__device__ __forceinline__ float ang_factor(const AngParams& A, int b, int fy, int fx) {
    const float xv = (float)fx * A.inv_hw, yv = (float)fy * A.inv_hh;
    const float r2 = xv * xv + yv * yv;
    float c = (r2 > 0.f) ? (xv * A.cs[b] + yv * A.sn[b]) * rsqrtf(r2) : A.cs[b];  // angle(DC) = atan2(0,0) = 0
    if (A.one_sided && !(c > 0.f)) return 0.f;
    float v = A.scale;
    for (int i = 0; i < A.order; ++i) v *= c;
    return v;
}

struct Tile {  // carve dynamic smem
    float2 *a, *b, *c;
};

// ------------------------------------------------------------------------------------------------
// K1: row pass, forward, with input prologue.  in -> FFT along x -> T[n][b][y][kx]
//   MODE 0: real input  in0[n][y][x]
//   MODE 1: polar input phase=in0, amp=in1 at channel (n*nbB + b)   (values_to_coeff, pyramid.py:103-108)
//   MODE 2: complex interleaved in0[(b)][n][y][x][2] via pointer table (band tensors [N,h,w,2])
// grid: (ceil(h/RB), nbB, N)
// ------------------------------------------------------------------------------------------------
struct PtrTable { const float* p[MAX_BANDS]; };

template <int MODE>
__global__ void __launch_bounds__(256) k_rows_fwd(Fft1D P, int h, int w, int RB, int nbB, const float* __restrict__ in0,
                                                  const float* __restrict__ in1, PtrTable tab,
                                                  float2* __restrict__ T) {
    extern __shared__ float2 smem[];
    float2* a = smem;
    float2* bq = smem + RB * w;
    const int y0 = blockIdx.x * RB, b = blockIdx.y, n = blockIdx.z, N = gridDim.z;
    const int rows = min(RB, h - y0);
    for (int q = threadIdx.x; q < rows * w; q += blockDim.x) {
        const int r = q / w, x = q - r * w;
        const size_t pix = (size_t)(y0 + r) * w + x;
        float2 z;
        if (MODE == 0) {
            z = make_float2(in0[(size_t)n * h * w + pix], 0.f);
        } else if (MODE == 1) {
            const size_t o = ((size_t)n * nbB + b) * h * w + pix;
            const float ph = in0[o], am = in1[o];
            float sn, cs;
            sincosf(ph, &sn, &cs);
            z = make_float2(cs * am, sn * am);  // pyramid.py:105-106
        } else {
            const float2* src = (const float2*)tab.p[b];
            z = src[(size_t)n * h * w + pix];
        }
        a[r * w + x] = z;
    }
    __syncthreads();
    float2* res = fft_smem<false>(P, a, bq, rows, w, 1);
    float2* dst = T + (((size_t)n * nbB + b) * h + y0) * w;
    for (int q = threadIdx.x; q < rows * w; q += blockDim.x) dst[q] = res[q];
    (void)N;
}

// ------------------------------------------------------------------------------------------------
// K2: column pass, forward, + combine (reconstruction) or plain in-place (decomposition of X).
//   For each band b: FFT along y of T[n][b][:, tile]; acc += fft * ang_b * fac ; out = acc * radial
//   plain (nbB==1, A==nullptr-like flag): out = fft * radial (radial may be null -> 1)
// grid: (ceil(w/CT), N)
// ------------------------------------------------------------------------------------------------
template <bool ANG>
__global__ void __launch_bounds__(512) k_cols_fwd(Fft1D P, int h, int w, int CT, int nbB, const float2* __restrict__ T,
                                                  const float* __restrict__ radial, AngParams A,
                                                  float2* __restrict__ out, size_t out_plane_stride) {
    extern __shared__ float2 smem[];
    const int E = h * CT;
    float2* a = smem;
    float2* bq = smem + E;
    float2* acc = smem + 2 * E;
    const int x0 = blockIdx.x * CT, n = blockIdx.y;
    const int cols = min(CT, w - x0);
    for (int b = 0; b < nbB; ++b) {
        const float2* src = T + ((size_t)n * nbB + b) * h * w;
        for (int q = threadIdx.x; q < E; q += blockDim.x) {
            const int y = q / CT, c = q - y * CT;
            a[q] = (c < cols) ? src[(size_t)y * w + x0 + c] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        float2* res = fft_smem<false>(P, a, bq, CT, 1, CT);
        if (ANG) {
            for (int q = threadIdx.x; q < E; q += blockDim.x) {
                const int ky = q / CT, c = q - ky * CT;
                const float g = ang_factor(A, b, sfreq(ky, h), sfreq(min(x0 + c, w - 1), w));
                const float2 v = cmul(make_float2(res[q].x * g, res[q].y * g), A.fac);
                acc[q] = (b == 0) ? v : cadd(acc[q], v);
            }
        } else {
            for (int q = threadIdx.x; q < E; q += blockDim.x) acc[q] = res[q];
        }
        __syncthreads();
    }
    float2* dst = out + (size_t)n * out_plane_stride;
    for (int q = threadIdx.x; q < E; q += blockDim.x) {
        const int ky = q / CT, c = q - ky * CT;
        if (c >= cols) continue;
        const size_t o = (size_t)ky * w + x0 + c;
        const float m = radial ? __ldg(radial + o) : 1.f;
        dst[o] = make_float2(acc[q].x * m, acc[q].y * m);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: column pass, inverse, with spectrum loader.
//   GATHER == false (decomposition): value = X[n][fy mod H][fx mod W] * radial[ky][kx] * ang_b * fac
//   GATHER == true  (reconstruction): value = sum over levels of Y_l[n][fy mod h_l][fx mod w_l] (+ high spectrum)
// out: T[n][b][y][kx]     grid: (ceil(w/CT), N)
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
    int nlev;                     // number of level spectra (band levels + low)
    LevelGeom lv[MAX_LEVELS];
    unsigned long long active;    // bit l set = level l contributes
    const float2* Y;              // region C base
    size_t plane_stride;          // complex elements per plane in region C
    const float2* Yhigh;          // [N][H][W] or null
};

template <bool ANG>
__global__ void __launch_bounds__(512) k_cols_inv_decomp(Fft1D P, int h, int w, int H, int W, int CT, int nbB,
                                                         const float2* __restrict__ X, const float* __restrict__ radial,
                                                         AngParams A, float2* __restrict__ T) {
    extern __shared__ float2 smem[];
    const int E = h * CT;
    float2* a = smem;
    float2* bq = smem + E;
    const int x0 = blockIdx.x * CT, n = blockIdx.y;
    const int cols = min(CT, w - x0);
    const float2* Xn = X + (size_t)n * H * W;
    for (int b = 0; b < nbB; ++b) {
        for (int q = threadIdx.x; q < E; q += blockDim.x) {
            const int ky = q / CT, c = q - ky * CT;
            float2 v = make_float2(0.f, 0.f);
            if (c < cols) {
                const int kx = x0 + c;
                const int fy = sfreq(ky, h), fx = sfreq(kx, w);
                float m = __ldg(radial + (size_t)ky * w + kx);
                if (ANG && m != 0.f) m *= ang_factor(A, b, fy, fx);
                if (m != 0.f) {
                    const float2 xv = __ldg(Xn + (size_t)wrapi(fy, H) * W + wrapi(fx, W));
                    v = make_float2(xv.x * m, xv.y * m);
                    if (ANG) v = cmul(v, A.fac);
                }
            }
            a[q] = v;
        }
        __syncthreads();
        float2* res = fft_smem<true>(P, a, bq, CT, 1, CT);
        float2* dst = T + ((size_t)n * nbB + b) * h * w;
        for (int q = threadIdx.x; q < E; q += blockDim.x) {
            const int y = q / CT, c = q - y * CT;
            if (c < cols) dst[(size_t)y * w + x0 + c] = res[q];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(512) k_cols_inv_gather(Fft1D P, int H, int W, int CT, GatherArgs G, float inv_hh,
                                                         float inv_hw, float2* __restrict__ T) {
    extern __shared__ float2 smem[];
    const int E = H * CT;
    float2* a = smem;
    float2* bq = smem + E;
    const int x0 = blockIdx.x * CT, n = blockIdx.y;
    const int cols = min(CT, W - x0);
    for (int q = threadIdx.x; q < E; q += blockDim.x) {
        const int ky = q / CT, c = q - ky * CT;
        float2 v = make_float2(0.f, 0.f);
        if (c < cols) {
            const int kx = x0 + c;
            const int fy = sfreq(ky, H), fx = sfreq(kx, W);
            const float xv = (float)fx * inv_hw, yv = (float)fy * inv_hh;
            const float r2 = xv * xv + yv * yv;
            if (G.Yhigh) v = G.Yhigh[((size_t)n * H + ky) * W + kx];
            for (int l = 0; l < G.nlev; ++l) {
                const LevelGeom g = G.lv[l];
                // nested centred windows: once outside, outside of all coarser levels too
                if (fy < -(g.h >> 1) || fy > g.h - 1 - (g.h >> 1) || fx < -(g.w >> 1) || fx > g.w - 1 - (g.w >> 1)) break;
                if (!((G.active >> l) & 1ull)) continue;
                const bool dc = (fy == 0 && fx == 0);
                if (!dc && (r2 <= g.rad2_lo || r2 >= g.rad2_hi)) continue;
                const float2 y = G.Y[(size_t)n * G.plane_stride + g.off + (size_t)wrapi(fy, g.h) * g.w + wrapi(fx, g.w)];
                v = cadd(v, y);
            }
        }
        a[q] = v;
    }
    __syncthreads();
    float2* res = fft_smem<true>(P, a, bq, CT, 1, CT);
    float2* dst = T + (size_t)n * H * W;
    for (int q = threadIdx.x; q < E; q += blockDim.x) {
        const int y = q / CT, c = q - y * CT;
        if (c < cols) dst[(size_t)y * W + x0 + c] = res[q];
    }
}

// ------------------------------------------------------------------------------------------------
// K4: row pass, inverse, with output epilogue.  T[n][b][y][:] -> IFFT along x -> * scale ->
//   EPI 0: real part -> out0[n][y][x]
//   EPI 1: polar: phase -> out0, amplitude -> out1 at channel n*nbB + b; per-plane max amplitude
//          (pyramid.py:63-69, phase_net.py:47-59)
//   EPI 2: complex interleaved -> tab.p[b][n][y][x][2]
// grid: (ceil(h/RB), nbB, N)
// ------------------------------------------------------------------------------------------------
struct MutPtrTable { float* p[MAX_BANDS]; };

template <int EPI>
__global__ void __launch_bounds__(256) k_rows_inv(Fft1D P, int h, int w, int RB, int nbB, const float2* __restrict__ T,
                                                  float scale, float* __restrict__ out0, float* __restrict__ out1,
                                                  MutPtrTable tab, float* __restrict__ amp_max) {
    extern __shared__ float2 smem[];
    float2* a = smem;
    float2* bq = smem + RB * w;
    const int y0 = blockIdx.x * RB, b = blockIdx.y, n = blockIdx.z;
    const int rows = min(RB, h - y0);
    const float2* src = T + (((size_t)n * nbB + b) * h + y0) * w;
    for (int q = threadIdx.x; q < rows * w; q += blockDim.x) a[q] = src[q];
    __syncthreads();
    float2* res = fft_smem<true>(P, a, bq, rows, w, 1);
    float mx = 0.f;
    for (int q = threadIdx.x; q < rows * w; q += blockDim.x) {
        const float2 z = make_float2(res[q].x * scale, res[q].y * scale);
        const size_t pix = (size_t)y0 * w + q;
        if (EPI == 0) {
            out0[(size_t)n * h * w + pix] = z.x;
        } else if (EPI == 1) {
            const size_t o = ((size_t)n * nbB + b) * h * w + pix;
            const float am = sqrtf(z.x * z.x + z.y * z.y);       // torch.abs            (pyramid.py:67)
            out0[o] = atan2f(z.y, z.x);                          // imag(log z)          (pyramid.py:63)
            out1[o] = am;
            mx = fmaxf(mx, am);
        } else {
            ((float2*)tab.p[b])[(size_t)n * h * w + pix] = z;
        }
    }
    if (EPI == 1 && amp_max) {
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) atomicMax((int*)(amp_max + n), __float_as_int(mx));  // amplitudes are >= 0
    }
}

// ------------------------------------------------------------------------------------------------
// host: launch helpers
// ------------------------------------------------------------------------------------------------
static int pick_rb(int h, int w) { return std::max(1, std::min(std::min(h, 32), ROW_ELEMS / w)); }
static int pick_ct(int h, int w, int nbuf_elems) {
    int ct = 32;
    while (ct > 1 && (size_t)ct * h > (size_t)nbuf_elems) ct >>= 1;
    while (ct > 1 && ct >= 2 * w) ct >>= 1;
    return ct;
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) { set_error("pyramid: tile needs %zu B of shared memory (> 227 KB)", bytes); return FVFI_EINVAL; }
    if (bytes > 48 * 1024) FVFI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FVFI_OK;
}

template <int MODE>
static int launch_rows_fwd(const Fft1D& P, int h, int w, int nbB, int N, const float* in0, const float* in1,
                           const PtrTable& tab, float2* T, cudaStream_t s) {
    const int RB = pick_rb(h, w);
    const size_t smem = (size_t)2 * RB * w * sizeof(float2);
    if (int rc = ensure_smem(k_rows_fwd<MODE>, smem)) return rc;
    dim3 grid(ceil_div(h, RB), nbB, N);
    k_rows_fwd<MODE><<<grid, 256, smem, s>>>(P, h, w, RB, nbB, in0, in1, tab, T);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

template <bool ANG>
static int launch_cols_fwd(const Fft1D& P, int h, int w, int nbB, int N, const float2* T, const float* radial,
                           const AngParams& A, float2* out, size_t stride, cudaStream_t s) {
    const int CT = pick_ct(h, w, COL_ELEMS);
    const size_t smem = (size_t)3 * h * CT * sizeof(float2);
    if (int rc = ensure_smem(k_cols_fwd<ANG>, smem)) return rc;
    dim3 grid(ceil_div(w, CT), N);
    k_cols_fwd<ANG><<<grid, (h * CT > 2048) ? 512 : 256, smem, s>>>(P, h, w, CT, nbB, T, radial, A, out, stride);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

template <bool ANG>
static int launch_cols_inv_decomp(const Fft1D& P, int h, int w, int H, int W, int nbB, int N, const float2* X,
                                  const float* radial, const AngParams& A, float2* T, cudaStream_t s) {
    const int CT = pick_ct(h, w, COL_ELEMS);
    const size_t smem = (size_t)2 * h * CT * sizeof(float2);
    if (int rc = ensure_smem(k_cols_inv_decomp<ANG>, smem)) return rc;
    dim3 grid(ceil_div(w, CT), N);
    k_cols_inv_decomp<ANG><<<grid, (h * CT > 2048) ? 512 : 256, smem, s>>>(P, h, w, H, W, CT, nbB, X, radial, A, T);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

template <int EPI>
static int launch_rows_inv(const Fft1D& P, int h, int w, int nbB, int N, const float2* T, float scale, float* out0,
                           float* out1, const MutPtrTable& tab, float* amp_max, cudaStream_t s) {
    const int RB = pick_rb(h, w);
    const size_t smem = (size_t)2 * RB * w * sizeof(float2);
    if (int rc = ensure_smem(k_rows_inv<EPI>, smem)) return rc;
    dim3 grid(ceil_div(h, RB), nbB, N);
    k_rows_inv<EPI><<<grid, 256, smem, s>>>(P, h, w, RB, nbB, T, scale, out0, out1, tab, amp_max);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

struct Workspace {
    float2 *A, *B, *C;  // A: N*H*W ; B: N*nb*H*W ; C: N*level_elems
};

static size_t ws_elems(const fvfi_pyr_plan* p, int N) {
    const size_t HW = (size_t)p->H * p->W;
    return (size_t)N * (HW + (size_t)p->nbands * HW + p->level_elems) + 64;
}

static Workspace carve(const fvfi_pyr_plan* p, int N, void* ws) {
    const size_t HW = (size_t)p->H * p->W;
    Workspace w;
    uintptr_t base = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    w.A = (float2*)base;
    w.B = w.A + (size_t)N * HW;
    w.C = w.B + (size_t)N * p->nbands * HW;
    return w;
}

// decomposition shared by the polar and complex front ends
static int decompose(const fvfi_pyr_plan* p, const float* img, int N, float* high, float* const* phase,
                     float* const* amp, float* const* bands, float* low, float* amp_max, void* workspace,
                     cudaStream_t s) {
    const int H = p->H, W = p->W, L = p->L, nb = p->nbands;
    Workspace ws = carve(p, N, workspace);
    PtrTable none{};
    MutPtrTable mnone{};
    AngParams noang{};
    // X = FFT2(img): rows then columns (in place in region A)
    if (int rc = launch_rows_fwd<0>(p->fx[0], H, W, 1, N, img, nullptr, none, ws.B, s)) return rc;
    if (int rc = launch_cols_fwd<false>(p->fy[0], H, W, 1, N, ws.B, nullptr, noang, ws.A, (size_t)H * W, s)) return rc;
    if (amp_max) FVFI_CUDA(cudaMemsetAsync(amp_max, 0, (size_t)L * N * sizeof(float), s));
    // high-pass residual
    if (high) {
        if (int rc = launch_cols_inv_decomp<false>(p->fy[0], H, W, H, W, 1, N, ws.A, p->hi0, noang, ws.B, s)) return rc;
        if (int rc = launch_rows_inv<0>(p->fx[0], H, W, 1, N, ws.B, 1.f / ((float)H * W), high, nullptr, mnone, nullptr, s))
            return rc;
    }
    // oriented band-pass levels
    for (int l = 0; l < L; ++l) {
        const int h = p->lv[l].h, w = p->lv[l].w;
        if (phase ? (phase[l] == nullptr) : (bands[l * nb] == nullptr)) continue;
        if (int rc = launch_cols_inv_decomp<true>(p->fy[l], h, w, H, W, nb, N, ws.A, p->radial[l], p->ang_build, ws.B, s))
            return rc;
        const float scale = 1.f / ((float)h * w);
        if (phase) {
            if (int rc = launch_rows_inv<1>(p->fx[l], h, w, nb, N, ws.B, scale, phase[l], amp[l], mnone,
                                            amp_max ? amp_max + (size_t)l * N : nullptr, s))
                return rc;
        } else {
            MutPtrTable t{};
            for (int b = 0; b < nb; ++b) t.p[b] = bands[l * nb + b];
            if (int rc = launch_rows_inv<2>(p->fx[l], h, w, nb, N, ws.B, scale, nullptr, nullptr, t, nullptr, s)) return rc;
        }
    }
    // low-pass residual
    if (low) {
        const int h = p->lv[L].h, w = p->lv[L].w;
        if (int rc = launch_cols_inv_decomp<false>(p->fy[L], h, w, H, W, 1, N, ws.A, p->radial[L], noang, ws.B, s)) return rc;
        if (int rc = launch_rows_inv<0>(p->fx[L], h, w, 1, N, ws.B, 1.f / ((float)h * w), low, nullptr, mnone, nullptr, s))
            return rc;
    }
    return FVFI_OK;
}

static int reconstruct(const fvfi_pyr_plan* p, const float* high, const float* const* phase, const float* const* amp,
                       const float* const* bands, const float* low, int N, float* img, void* workspace,
                       cudaStream_t s) {
    const int H = p->H, W = p->W, L = p->L, nb = p->nbands;
    Workspace ws = carve(p, N, workspace);
    PtrTable none{};
    MutPtrTable mnone{};
    AngParams noang{};
    GatherArgs G{};
    G.nlev = L + 1;
    G.Y = ws.C;
    G.plane_stride = p->level_elems;
    G.active = 0;
    for (int l = 0; l <= L; ++l) G.lv[l] = p->lv[l];
    for (int l = 0; l < L; ++l) {
        const int h = p->lv[l].h, w = p->lv[l].w;
        const bool have = phase ? (phase[l] != nullptr && amp[l] != nullptr) : (bands[l * nb] != nullptr);
        if (!have) continue;
        if (phase) {
            if (int rc = launch_rows_fwd<1>(p->fx[l], h, w, nb, N, phase[l], amp[l], none, ws.B, s)) return rc;
        } else {
            PtrTable t{};
            for (int b = 0; b < nb; ++b) t.p[b] = bands[l * nb + b];
            if (int rc = launch_rows_fwd<2>(p->fx[l], h, w, nb, N, nullptr, nullptr, t, ws.B, s)) return rc;
        }
        if (int rc = launch_cols_fwd<true>(p->fy[l], h, w, nb, N, ws.B, p->radial[l], p->ang_rec, ws.C + p->lv[l].off,
                                           p->level_elems, s))
            return rc;
        G.active |= 1ull << l;
    }
    if (low) {
        const int h = p->lv[L].h, w = p->lv[L].w;
        if (int rc = launch_rows_fwd<0>(p->fx[L], h, w, 1, N, low, nullptr, none, ws.B, s)) return rc;
        if (int rc = launch_cols_fwd<false>(p->fy[L], h, w, 1, N, ws.B, p->radial[L], noang, ws.C + p->lv[L].off,
                                            p->level_elems, s))
            return rc;
        G.active |= 1ull << L;
    }
    if (high) {
        if (int rc = launch_rows_fwd<0>(p->fx[0], H, W, 1, N, high, nullptr, none, ws.B, s)) return rc;
        if (int rc = launch_cols_fwd<false>(p->fy[0], H, W, 1, N, ws.B, p->hi0, noang, ws.A, (size_t)H * W, s)) return rc;
        G.Yhigh = ws.A;
    }
    // gather all level spectra + inverse FFT2
    {
        const int CT = pick_ct(H, W, COL_ELEMS);
        const size_t smem = (size_t)2 * H * CT * sizeof(float2);
        if (int rc = ensure_smem(k_cols_inv_gather, smem)) return rc;
        dim3 grid(ceil_div(W, CT), N);
        k_cols_inv_gather<<<grid, 512, smem, s>>>(p->fy[0], H, W, CT, G, p->ang_rec.inv_hh, p->ang_rec.inv_hw, ws.B);
        FVFI_LAUNCH_CHECK();
    }
    return launch_rows_inv<0>(p->fx[0], H, W, 1, N, ws.B, 1.f / ((float)H * W), img, nullptr, mnone, nullptr, s);
}

}  // namespace fvfi

using namespace fvfi;

extern "C" {

int fvfi_pyr_next_size(int n, double s) { return fvfi::next_size(n, s); }

int fvfi_pyr_plan_create(int H, int W, int height, int nbands, double scale_factor, fvfi_pyr_plan** out) {
    FVFI_CHECK_ARG(out, "pyr_plan_create: null out");
    *out = nullptr;
    FVFI_CHECK_ARG(H >= 4 && W >= 4 && H <= 8192 && W <= 8192, "pyr_plan_create: unsupported image size %dx%d", H, W);
    FVFI_CHECK_ARG(height >= 2 && height - 2 < MAX_LEVELS - 1, "pyr_plan_create: bad height %d", height);
    FVFI_CHECK_ARG(nbands >= 1 && nbands <= MAX_BANDS, "pyr_plan_create: nbands must be 1..%d", MAX_BANDS);
    FVFI_CHECK_ARG(scale_factor > 1.0 && scale_factor <= 4.0, "pyr_plan_create: scale_factor must be in (1,4]");
    fvfi_pyr_plan* p = new fvfi_pyr_plan();
    p->H = H; p->W = W; p->height = height; p->nbands = nbands; p->L = height - 2; p->scale = scale_factor;
    if (int rc = build_plan(p)) { fvfi_pyr_plan_destroy(p); return rc; }
    *out = p;
    return FVFI_OK;
}

void fvfi_pyr_plan_destroy(fvfi_pyr_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    delete p;
}

int fvfi_pyr_num_levels(const fvfi_pyr_plan* p) { return p ? p->L : -1; }

int fvfi_pyr_level_shape(const fvfi_pyr_plan* p, int level, int* h, int* w) {
    FVFI_CHECK_ARG(p && level >= 0 && level <= p->L, "pyr_level_shape: bad level");
    if (h) *h = p->lv[level].h;
    if (w) *w = p->lv[level].w;
    return FVFI_OK;
}

size_t fvfi_pyr_workspace_bytes(const fvfi_pyr_plan* p, int N) {
    if (!p || N <= 0) return 0;
    return ws_elems(p, N) * sizeof(float2) + 256;
}

int fvfi_pyr_decompose(const fvfi_pyr_plan* p, const float* img, int N, float* high, float* const* phase,
                       float* const* amp, float* low, float* amp_max, void* workspace, void* stream) {
    FVFI_CHECK_ARG(p && img && phase && amp && workspace && N > 0 && N <= 65535, "pyr_decompose: bad argument");
    return decompose(p, img, N, high, phase, amp, nullptr, low, amp_max, workspace, (cudaStream_t)stream);
}

int fvfi_pyr_build_complex(const fvfi_pyr_plan* p, const float* img, int N, float* high, float* const* bands,
                           float* low, void* workspace, void* stream) {
    FVFI_CHECK_ARG(p && img && bands && workspace && N > 0 && N <= 65535, "pyr_build_complex: bad argument");
    return decompose(p, img, N, high, nullptr, nullptr, bands, low, nullptr, workspace, (cudaStream_t)stream);
}

int fvfi_pyr_reconstruct(const fvfi_pyr_plan* p, const float* high, const float* const* phase,
                         const float* const* amp, const float* low, int N, float* img, void* workspace,
                         void* stream) {
    FVFI_CHECK_ARG(p && phase && amp && img && workspace && N > 0 && N <= 65535, "pyr_reconstruct: bad argument");
    return reconstruct(p, high, phase, amp, nullptr, low, N, img, workspace, (cudaStream_t)stream);
}

int fvfi_pyr_reconstruct_complex(const fvfi_pyr_plan* p, const float* high, const float* const* bands,
                                 const float* low, int N, float* img, void* workspace, void* stream) {
    FVFI_CHECK_ARG(p && bands && img && workspace && N > 0 && N <= 65535, "pyr_reconstruct_complex: bad argument");
    return reconstruct(p, high, nullptr, nullptr, bands, low, N, img, workspace, (cudaStream_t)stream);
}

}  // extern "C"
