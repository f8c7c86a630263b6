// adacof_tiled.cu -- smem-staged AdaCoF warp (forward, fused backward, fused two-frame synthesis).
//
// Why: with one thread per output pixel the 4*F*F*C frame taps per pixel are data-dependent
// gathers.  Through L1 a warp-wide gather costs one wavefront per distinct 128 B line (up to 32
// when the per-pixel offsets are uncorrelated), which makes the reference kernel -- and our
// "direct" family -- L1-wavefront-bound far below the HBM roofline.  Here a CTA stages the
// frame region its output tile can reach (tile + (F-1)*dilation + a halo of R pixels each
// side) into shared memory as channel-interleaved float4 {R,G,B,0} pixels, so one tap of all
// three channels is ONE conflict-light LDS.128, and the (r0,c0),(r0,c0+1),(r0+1,c0),(r0+1,c0+1)
// neighbours are immediate offsets from one address.  The staged region is filled with
// CLAMPED coordinates, so the reference's clamp-to-edge (adacof.py:30-52) is implicit for
// in-region taps.  Taps whose offset leaves the region (|offset| > R) take a per-lane global
// fallback with the explicit clamps; semantics are identical for any offset.
//
// The coefficient maps (3*F*F planes, 300 B/pixel at F=5: the HBM-dominant stream) are read
// once, coalesced (lane <-> consecutive pixel), L1-bypassing, one tap-row (3*F values) ahead
// of use to keep ~60 KB/SM in flight.
#include "common.cuh"

namespace fvfi {

constexpr int TW = 64;       // output tile width  (2 pixels per lane per row)
constexpr int TH = 16;       // output tile height
constexpr int NTHREADS = 256;
constexpr int HALO = 8;      // offsets with -(HALO+1) < offset < HALO+1 stay in the staged region
constexpr int MINB = 3;      // resident CTAs per SM the register budget is tuned for

template <int PADMAX>
struct Region {
    static constexpr int SH = TH + PADMAX + 2 * HALO + 1;
    static constexpr int SW = TW + PADMAX + 2 * HALO + 1;
    static constexpr size_t BYTES = (size_t)SH * SW * sizeof(float4);
};

// Stage frame region (clamped coordinates) as float4 pixels.
template <int PADMAX>
__device__ __forceinline__ void stage_region(float4* __restrict__ sm, const float* __restrict__ I,
                                             size_t plane_in, int Hin, int Win, int i0, int j0) {
    using Rg = Region<PADMAX>;
    for (int p = threadIdx.x; p < Rg::SH * Rg::SW; p += NTHREADS) {
        const int r = p / Rg::SW, c = p - r * Rg::SW;
        const int gr = min(max(i0 - HALO + r, 0), Hin - 1);
        const int gc = min(max(j0 - HALO + c, 0), Win - 1);
        const float* src = I + (size_t)gr * Win + gc;
        float4 v;
        v.x = __ldg(src);
        v.y = __ldg(src + plane_in);
        v.z = __ldg(src + 2 * plane_in);
        v.w = 0.f;
        sm[p] = v;
    }
}

struct Quad {
    float4 v00, v10, v01, v11;
};

// Gather the four neighbours of one tap: in-region -> LDS.128 x4; else global with clamps.
template <int PADMAX>
__device__ __forceinline__ Quad gather(const float4* __restrict__ sm, const float* __restrict__ I,
                                       size_t plane_in, int Hin, int Win, int rr, int cc, int gr, int gc) {
    using Rg = Region<PADMAX>;
    Quad q;
    if ((unsigned)rr < (unsigned)(Rg::SH - 1) && (unsigned)cc < (unsigned)(Rg::SW - 1)) {
        const float4* p = sm + rr * Rg::SW + cc;
        q.v00 = p[0];
        q.v01 = p[1];
        q.v10 = p[Rg::SW];
        q.v11 = p[Rg::SW + 1];
    } else {
        const int r0 = min(max(gr, 0), Hin - 1), r1 = min(max(gr + 1, 0), Hin - 1);
        const int c0 = min(max(gc, 0), Win - 1), c1 = min(max(gc + 1, 0), Win - 1);
        const float* p00 = I + (size_t)r0 * Win + c0;
        const float* p10 = I + (size_t)r1 * Win + c0;
        const float* p01 = I + (size_t)r0 * Win + c1;
        const float* p11 = I + (size_t)r1 * Win + c1;
        q.v00 = make_float4(__ldg(p00), __ldg(p00 + plane_in), __ldg(p00 + 2 * plane_in), 0.f);
        q.v10 = make_float4(__ldg(p10), __ldg(p10 + plane_in), __ldg(p10 + 2 * plane_in), 0.f);
        q.v01 = make_float4(__ldg(p01), __ldg(p01 + plane_in), __ldg(p01 + 2 * plane_in), 0.f);
        q.v11 = make_float4(__ldg(p11), __ldg(p11 + plane_in), __ldg(p11 + 2 * plane_in), 0.f);
    }
    return q;
}

struct Moments {
    float s0, s1i, s2i, s1j, s2j;
};

// One pixel, all taps, forward accumulate.  FT > 0: compile-time filter size (fully unrolled,
// coefficient loads issued one tap-row ahead); FT == 0: runtime F.
template <int FT, int PADMAX, bool MOMENTS>
__device__ __forceinline__ void warp_pixel(const float4* __restrict__ sm, const float* __restrict__ I,
                                           size_t plane_in, int Hin, int Win, const float* __restrict__ wq,
                                           const float* __restrict__ aq, const float* __restrict__ bq,
                                           size_t plane, int F, int dil, int li, int lj, int gi, int gj,
                                           float acc[3], Moments& m) {
    acc[0] = acc[1] = acc[2] = 0.f;
    if (MOMENTS) m.s0 = m.s1i = m.s2i = m.s1j = m.s2j = 0.f;
    auto tap = [&](float w, float al, float be, int k, int l) {
        const int A = (int)al, B = (int)be;               // trunc toward zero -- adacof.py:27-28
        const float a = al - (float)A, b = be - (float)B;
        const float na = 1.f - a, nb = 1.f - b;
        const Quad q = gather<PADMAX>(sm, I, plane_in, Hin, Win, li + k * dil + A + HALO, lj + l * dil + B + HALO,
                                      gi + k * dil + A, gj + l * dil + B);
        const float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
        acc[0] = fmaf(w, q.v00.x * w00 + q.v10.x * w10 + q.v01.x * w01 + q.v11.x * w11, acc[0]);
        acc[1] = fmaf(w, q.v00.y * w00 + q.v10.y * w10 + q.v01.y * w01 + q.v11.y * w11, acc[1]);
        acc[2] = fmaf(w, q.v00.z * w00 + q.v10.z * w10 + q.v01.z * w01 + q.v11.z * w11, acc[2]);
        if (MOMENTS) {
            m.s0 += w;
            m.s1i = fmaf(w, al, m.s1i);
            m.s2i = fmaf(w * al, al, m.s2i);
            m.s1j = fmaf(w, be, m.s1j);
            m.s2j = fmaf(w * be, be, m.s2j);
        }
    };
    if (FT > 0) {
        float w[FT > 0 ? FT : 1], al[FT > 0 ? FT : 1], be[FT > 0 ? FT : 1];
#pragma unroll
        for (int l = 0; l < FT; ++l) {
            w[l] = ld_stream(wq + (size_t)l * plane);
            al[l] = ld_stream(aq + (size_t)l * plane);
            be[l] = ld_stream(bq + (size_t)l * plane);
        }
#pragma unroll
        for (int k = 0; k < FT; ++k) {
            float wn[FT > 0 ? FT : 1], an[FT > 0 ? FT : 1], bn[FT > 0 ? FT : 1];
            if (k + 1 < FT) {
#pragma unroll
                for (int l = 0; l < FT; ++l) {
                    const size_t o = (size_t)((k + 1) * FT + l) * plane;
                    wn[l] = ld_stream(wq + o);
                    an[l] = ld_stream(aq + o);
                    bn[l] = ld_stream(bq + o);
                }
            }
#pragma unroll
            for (int l = 0; l < FT; ++l) tap(w[l], al[l], be[l], k, l);
            if (k + 1 < FT) {
#pragma unroll
                for (int l = 0; l < FT; ++l) {
                    w[l] = wn[l];
                    al[l] = an[l];
                    be[l] = bn[l];
                }
            }
        }
    } else {
        size_t o = 0;
        for (int k = 0; k < F; ++k)
            for (int l = 0; l < F; ++l, o += plane) tap(ld_stream(wq + o), ld_stream(aq + o), ld_stream(bq + o), k, l);
    }
}

// ---------------------------------------------------------------------------------------------
// forward (NFRAMES == 1) and fused AdaCoFNet synthesis (NFRAMES == 2: t1, t2, blend, mask)
// ---------------------------------------------------------------------------------------------
struct FwdArgs {
    const float* in[2];
    const float* w[2];
    const float* a[2];
    const float* b[2];
    const float* occ;
    float* t[2];
    float* frame;
    float* mask;
    int Hin, Win, H, W, F, dil;
};

template <int FT, int PADMAX, int NFRAMES>
__global__ void __launch_bounds__(NTHREADS, MINB) adacof_fwd_tiled(const FwdArgs A) {
    extern __shared__ float4 sm[];
    using Rg = Region<PADMAX>;
    constexpr int C = 3;
    const int j0 = blockIdx.x * TW, i0 = blockIdx.y * TH, n = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)A.H * A.W, plane_in = (size_t)A.Hin * A.Win;
    const int FF = A.F * A.F;
    // NFRAMES == 2: frame-1 results {r,g,b,var} wait in smem (behind the staged region) for the blend
    float4* stash = sm + Rg::SH * Rg::SW;

#pragma unroll
    for (int f = 0; f < NFRAMES; ++f) {
        const float* I = A.in[f] + (size_t)n * C * plane_in;
        if (f > 0) __syncthreads();  // everyone done reading the previous frame's region
        stage_region<PADMAX>(sm, I, plane_in, A.Hin, A.Win, i0, j0);
        __syncthreads();
#pragma unroll 1
        for (int px = 0; px < 4; ++px) {
            const int li = warp + (px >> 1) * 8, lj = lane + (px & 1) * 32;
            const int gi = i0 + li, gj = j0 + lj;
            if (gi >= A.H || gj >= A.W) continue;
            const size_t p = (size_t)gi * A.W + gj;
            const size_t q = (size_t)n * FF * plane + p;
            Moments m;
            float r[3];
            warp_pixel<FT, PADMAX, NFRAMES == 2>(sm, I, plane_in, A.Hin, A.Win, A.w[f] + q, A.a[f] + q, A.b[f] + q,
                                                 plane, A.F, A.dil, li, lj, gi, gj, r, m);
            float* const t = A.t[f];
            if (t) {
#pragma unroll
                for (int c = 0; c < C; ++c) st_stream(t + ((size_t)n * C + c) * plane + p, r[c]);
            }
            if (NFRAMES == 2) {
                const float var = (m.s2i - m.s1i * m.s1i * (2.f - m.s0)) + (m.s2j - m.s1j * m.s1j * (2.f - m.s0));
                if (f == 0) {
                    stash[px * NTHREADS + threadIdx.x] = make_float4(r[0], r[1], r[2], var);
                } else {
                    const float4 r1 = stash[px * NTHREADS + threadIdx.x];
                    if (A.frame) {  // fusion_adacofnet.py:198
                        const float o = ld_stream(A.occ + (size_t)n * plane + p);
                        st_stream(A.frame + ((size_t)n * C + 0) * plane + p, o * r1.x + (1.f - o) * r[0]);
                        st_stream(A.frame + ((size_t)n * C + 1) * plane + p, o * r1.y + (1.f - o) * r[1]);
                        st_stream(A.frame + ((size_t)n * C + 2) * plane + p, o * r1.z + (1.f - o) * r[2]);
                    }
                    if (A.mask)  // fusion_adacofnet.py:211-213
                        st_stream(A.mask + (size_t)n * plane + p, fminf(fmaxf(fmaxf(r1.w, var), 0.f), 20.f) / 20.f);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused backward: gW, g_alpha, g_beta (see adacof.cu for the algebra)
// ---------------------------------------------------------------------------------------------
template <int FT, int PADMAX>
__global__ void __launch_bounds__(NTHREADS, MINB)
adacof_bwd_tiled(const float* __restrict__ gout, const float* __restrict__ input, const float* __restrict__ weight,
                 const float* __restrict__ off_i, const float* __restrict__ off_j, float* __restrict__ gw,
                 float* __restrict__ goi, float* __restrict__ goj, int Hin, int Win, int H, int W, int F, int dil) {
    extern __shared__ float4 sm[];
    constexpr int C = 3;
    const int j0 = blockIdx.x * TW, i0 = blockIdx.y * TH, n = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)H * W, plane_in = (size_t)Hin * Win;
    const int FF = F * F;
    const float* I = input + (size_t)n * C * plane_in;
    stage_region<PADMAX>(sm, I, plane_in, Hin, Win, i0, j0);
    __syncthreads();
#pragma unroll 1
    for (int px = 0; px < 4; ++px) {
        const int li = warp + (px >> 1) * 8, lj = lane + (px & 1) * 32;
        const int gi = i0 + li, gj = j0 + lj;
        if (gi >= H || gj >= W) continue;
        const size_t p = (size_t)gi * W + gj;
        const float g0 = ld_stream(gout + ((size_t)n * C + 0) * plane + p);
        const float g1 = ld_stream(gout + ((size_t)n * C + 1) * plane + p);
        const float g2 = ld_stream(gout + ((size_t)n * C + 2) * plane + p);
        const size_t q = (size_t)n * FF * plane + p;
        auto tap = [&](float w, float al, float be, int k, int l, size_t o) {
            const int A = (int)al, B = (int)be;
            const float a = al - (float)A, b = be - (float)B;
            const float na = 1.f - a, nb = 1.f - b;
            const Quad v = gather<PADMAX>(sm, I, plane_in, Hin, Win, li + k * dil + A + HALO, lj + l * dil + B + HALO,
                                          gi + k * dil + A, gj + l * dil + B);
            const float s00 = fmaf(g2, v.v00.z, fmaf(g1, v.v00.y, g0 * v.v00.x));
            const float s10 = fmaf(g2, v.v10.z, fmaf(g1, v.v10.y, g0 * v.v10.x));
            const float s01 = fmaf(g2, v.v01.z, fmaf(g1, v.v01.y, g0 * v.v01.x));
            const float s11 = fmaf(g2, v.v11.z, fmaf(g1, v.v11.y, g0 * v.v11.x));
            st_stream(gw + q + o, s00 * (na * nb) + s10 * (a * nb) + s01 * (na * b) + s11 * (a * b));
            st_stream(goi + q + o, w * ((s10 - s00) * nb + (s11 - s01) * b));
            st_stream(goj + q + o, w * ((s01 - s00) * na + (s11 - s10) * a));
        };
        if (FT > 0) {
            float w[FT > 0 ? FT : 1], al[FT > 0 ? FT : 1], be[FT > 0 ? FT : 1];
#pragma unroll
            for (int l = 0; l < FT; ++l) {
                w[l] = ld_stream(weight + q + (size_t)l * plane);
                al[l] = ld_stream(off_i + q + (size_t)l * plane);
                be[l] = ld_stream(off_j + q + (size_t)l * plane);
            }
#pragma unroll
            for (int k = 0; k < FT; ++k) {
                float wn[FT > 0 ? FT : 1], an[FT > 0 ? FT : 1], bn[FT > 0 ? FT : 1];
                if (k + 1 < FT) {
#pragma unroll
                    for (int l = 0; l < FT; ++l) {
                        const size_t o = (size_t)((k + 1) * FT + l) * plane;
                        wn[l] = ld_stream(weight + q + o);
                        an[l] = ld_stream(off_i + q + o);
                        bn[l] = ld_stream(off_j + q + o);
                    }
                }
#pragma unroll
                for (int l = 0; l < FT; ++l) tap(w[l], al[l], be[l], k, l, (size_t)(k * FT + l) * plane);
                if (k + 1 < FT) {
#pragma unroll
                    for (int l = 0; l < FT; ++l) {
                        w[l] = wn[l];
                        al[l] = an[l];
                        be[l] = bn[l];
                    }
                }
            }
        } else {
            size_t o = 0;
            for (int k = 0; k < F; ++k)
                for (int l = 0; l < F; ++l, o += plane)
                    tap(ld_stream(weight + q + o), ld_stream(off_i + q + o), ld_stream(off_j + q + o), k, l, o);
        }
    }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    FVFI_SMEM_OPT_IN(kernel, bytes);
    return FVFI_OK;
}

template <int NFRAMES>
static int launch_fwd(const FwdArgs& a, int B, cudaStream_t s, int* handled) {
    const int pad = (a.F - 1) * a.dil;
    dim3 grid(ceil_div(a.W, TW), ceil_div(a.H, TH), B);
    const size_t stash_bytes = NFRAMES == 2 ? (size_t)4 * NTHREADS * sizeof(float4) : 0;
    *handled = 1;
    if (a.F == 5 && a.dil == 1) {
        auto k = adacof_fwd_tiled<5, 4, NFRAMES>;
        const size_t bytes = Region<4>::BYTES + stash_bytes;
        if (int rc = set_smem(k, bytes)) return rc;
        k<<<grid, NTHREADS, bytes, s>>>(a);
    } else if (pad <= 8) {
        auto k = adacof_fwd_tiled<0, 8, NFRAMES>;
        const size_t bytes = Region<8>::BYTES + stash_bytes;
        if (int rc = set_smem(k, bytes)) return rc;
        k<<<grid, NTHREADS, bytes, s>>>(a);
    } else {
        *handled = 0;
        return FVFI_OK;
    }
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

int adacof_forward_tiled(const float* input, const float* weight, const float* off_i, const float* off_j,
                         float* output, int B, int Hin, int Win, int H, int W, int F, int dil, cudaStream_t s,
                         int* handled) {
    FwdArgs a{};
    a.in[0] = input; a.w[0] = weight; a.a[0] = off_i; a.b[0] = off_j; a.t[0] = output;
    a.Hin = Hin; a.Win = Win; a.H = H; a.W = W; a.F = F; a.dil = dil;
    return launch_fwd<1>(a, B, s, handled);
}

int adacofnet_warp_blend_tiled(const float* in1, const float* in2, const float* w1, const float* a1,
                               const float* b1, const float* w2, const float* a2, const float* b2,
                               const float* occ, float* t1, float* t2, float* frame, float* mask, int B, int Hin,
                               int Win, int H, int W, int F, int dil, cudaStream_t s, int* handled) {
    FwdArgs a{};
    a.in[0] = in1; a.in[1] = in2; a.w[0] = w1; a.w[1] = w2; a.a[0] = a1; a.a[1] = a2; a.b[0] = b1; a.b[1] = b2;
    a.occ = occ; a.t[0] = t1; a.t[1] = t2; a.frame = frame; a.mask = mask;
    a.Hin = Hin; a.Win = Win; a.H = H; a.W = W; a.F = F; a.dil = dil;
    return launch_fwd<2>(a, B, s, handled);
}

int adacof_backward_tiled(const float* gout, const float* input, const float* weight, const float* off_i,
                          const float* off_j, float* gw, float* goi, float* goj, int B, int Hin, int Win, int H,
                          int W, int F, int dil, cudaStream_t s, int* handled) {
    const int pad = (F - 1) * dil;
    dim3 grid(ceil_div(W, TW), ceil_div(H, TH), B);
    *handled = 1;
    if (F == 5 && dil == 1) {
        auto k = adacof_bwd_tiled<5, 4>;
        if (int rc = set_smem(k, Region<4>::BYTES)) return rc;
        k<<<grid, NTHREADS, Region<4>::BYTES, s>>>(gout, input, weight, off_i, off_j, gw, goi, goj, Hin, Win, H, W, F,
                                                   dil);
    } else if (pad <= 8) {
        auto k = adacof_bwd_tiled<0, 8>;
        if (int rc = set_smem(k, Region<8>::BYTES)) return rc;
        k<<<grid, NTHREADS, Region<8>::BYTES, s>>>(gout, input, weight, off_i, off_j, gw, goi, goj, Hin, Win, H, W, F,
                                                   dil);
    } else {
        *handled = 0;
        return FVFI_OK;
    }
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

}  // namespace fvfi
