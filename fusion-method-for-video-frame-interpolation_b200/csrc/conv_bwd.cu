// conv_bwd.cu -- backward pass of the stride-1 "same" convolutions, for the FusionNet training step
// (src/fusion_net/trainer.py:246-259: L1 loss -> backward -> Adam; the reference gets these gradients from
// torch.nn.Conv2d -> cuDNN autograd for the seven layers of src/fusion_net/fusion_net.py:24-36).
//
//   y = act(conv(pad(x), w) + b),   g = dL/dy * act'(y)
//
//   fvfi_conv2d_grad_act          g, written into the interior of a zero canvas [B, H+2P, W+2P, C], and the bias gradient
//                                 gb[c] = sum_pixels g (deterministic: per-block partial sums, then one reduction).
//   data gradient (dgrad)         NOT a new kernel: dL/d(pad x) is the full correlation of g with the flipped, transposed filter
//                                 = the forward tcgen05 kernel (conv_tc.cu, 3xTF32: gradients have no bounded range) run over the
//                                 canvas with zero padding P; the host side (fvfi/conv.py) packs w[o,c,K-1-ky,K-1-kx] -> [c,o,ky,kx].
//   fvfi_reflect_pad_backward_nhwc   adjoint of torch's 'reflect' padding: folds the border of dL/d(pad x) back onto the image.
//   fvfi_conv2d_wgrad_nhwc        gw[o,c,ky,kx] = sum_{b,y,x} g[b,y,x,o] * padx[b,y+ky,x+kx,c]: a [Cout] x [K*K*Cin] GEMM whose
//                                 reduction runs over ALL pixels (5e5 at 8 crops of 256^2 against <= 2e5 outputs), so it is split
//                                 over pixel ranges (one CTA per output tile and range, fp32 FFMA, register-prefetched shared-memory
//                                 tiles, the im2col operand gathered with the padding rule on the fly) into partial sums that a
//                                 second kernel adds in a fixed order -> bit-reproducible gradients, which the data-parallel
//                                 parity check (N ranks == one process on the same samples) relies on.
//                                 CUDA cores on purpose: 62 GFLOP per training step in total, fp32-exact products, and the
//                                 tensor-core form would need both operands pixel-major (MN-major tf32), which tcgen05's shared-memory
//                                 descriptors do not offer without a transposing loader.
//   fvfi_max_pool2_backward_nhwc / fvfi_avg_pool2_backward_nhwc / fvfi_resize_bilinear_backward_nhwc / fvfi_fusion_blend_backward
//                                 adjoints of the other differentiable steps of FusionNet (src/fusion_net/fusion_net.py:52-77) and of
//                                 KernelEstimation's pooling / Upsample modules (src/fusion_net/fusion_adacofnet.py:29-36,62-70).
#include <algorithm>

#include "common.cuh"

namespace fvfi {

enum { BACT_NONE = 0, BACT_RELU = 1, BACT_ELU = 2, BACT_TANH = 3, BACT_SIGMOID = 4 };

// d act(v) / dv expressed through the saved OUTPUT y = act(v)
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
    switch (act) {
        case BACT_RELU: return y > 0.f ? 1.f : 0.f;
        case BACT_ELU: return y > 0.f ? 1.f : y + 1.f;      // elu'(v) = exp(v) = y + 1 for v < 0
        case BACT_TANH: return 1.f - y * y;
        case BACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// g = gy * act'(y) into the interior of a zero canvas; per-block column sums for the bias gradient
// block (CX, PY): thread (tx, ty) owns channels tx, tx + CX, ... of the canvas pixels ty, ty + PY, ... of the block's range
// ------------------------------------------------------------------------------------------------------------------
constexpr int GA_THREADS = 256;
constexpr int GA_MAX_CPT = 8;       // channel items per thread (VEC: float4 quads)

// One block walks canvas rows blockIdx.x, blockIdx.x + gridDim.x, ...; thread (tx, ty) owns the channel items tx, tx + CX, ... of the
// row's pixels ty, ty + PY, ...  VEC: an item is a float4 channel quad (C, both pixel strides multiples of 4, 16-byte aligned bases).
template <bool VEC>
__global__ void __launch_bounds__(GA_THREADS) grad_act_kernel(const float* __restrict__ gy, int gy_ps, const float* __restrict__ y,
                                                             int y_ps, float* __restrict__ canvas, int B, int H, int W, int C,
                                                             int border, int act, float* __restrict__ bias_part) {
    extern __shared__ float red[];                       // [PY][C]
    constexpr int E = VEC ? 4 : 1;
    const int CX = blockDim.x, PY = blockDim.y;
    const int Hc = H + 2 * border, Wc = W + 2 * border;
    const int items = C / E;
    float sum[GA_MAX_CPT][E];
#pragma unroll
    for (int i = 0; i < GA_MAX_CPT; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) sum[i][e] = 0.f;
    for (int row = blockIdx.x; row < B * Hc; row += gridDim.x) {
        const int b = row / Hc, yy = row - b * Hc - border;
        const bool row_in = yy >= 0 && yy < H;
        float* crow = canvas + (size_t)row * Wc * C;
        const size_t srow = ((size_t)b * H + (row_in ? yy : 0)) * W;
        for (int xc = threadIdx.y; xc < Wc; xc += PY) {
            const int xx = xc - border;
            const bool inside = row_in && xx >= 0 && xx < W;
            const size_t src = srow + (inside ? xx : 0);
#pragma unroll
            for (int i = 0; i < GA_MAX_CPT; ++i) {
                const int it = threadIdx.x + i * CX;
                if (it >= items) continue;
                float v[E];
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = 0.f;
                if (inside) {
                    if (VEC) {
                        const float4 gv = __ldg(reinterpret_cast<const float4*>(gy + src * gy_ps) + it);
                        v[0] = gv.x; v[E > 1 ? 1 : 0] = gv.y; v[E > 2 ? 2 : 0] = gv.z; v[E > 3 ? 3 : 0] = gv.w;
                        if (act) {
                            const float4 yv = __ldg(reinterpret_cast<const float4*>(y + src * y_ps) + it);
                            v[0] *= act_grad_from_output(yv.x, act);
                            v[E > 1 ? 1 : 0] *= act_grad_from_output(yv.y, act);
                            v[E > 2 ? 2 : 0] *= act_grad_from_output(yv.z, act);
                            v[E > 3 ? 3 : 0] *= act_grad_from_output(yv.w, act);
                        }
                    } else {
                        v[0] = __ldg(gy + src * gy_ps + it) * (act ? act_grad_from_output(__ldg(y + src * y_ps + it), act) : 1.f);
                    }
                }
                if (VEC) reinterpret_cast<float4*>(crow + (size_t)xc * C)[it] = make_float4(v[0], v[E > 1 ? 1 : 0], v[E > 2 ? 2 : 0], v[E > 3 ? 3 : 0]);
                else crow[(size_t)xc * C + it] = v[0];
#pragma unroll
                for (int e = 0; e < E; ++e) sum[i][e] += v[e];
            }
        }
    }
    if (bias_part == nullptr) return;
#pragma unroll
    for (int i = 0; i < GA_MAX_CPT; ++i) {
        const int it = threadIdx.x + i * CX;
        if (it < items)
#pragma unroll
            for (int e = 0; e < E; ++e) red[threadIdx.y * C + it * E + e] = sum[i][e];
    }
    __syncthreads();
    for (int c = threadIdx.y * CX + threadIdx.x; c < C; c += CX * PY) {
        float s = 0.f;
        for (int j = 0; j < PY; ++j) s += red[j * C + c];        // fixed order
        bias_part[(size_t)blockIdx.x * C + c] = s;
    }
}

// out[i] = sum_s part[s * n + i], fixed order; remap: wgrad partials are [Cout][tap][CinP], the gradient is OIHW [Cout][Cin][tap]
__global__ void sum_partials_kernel(const float* __restrict__ part, int S, size_t n, float* __restrict__ out, int Cin, int CinP, int taps) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t o = i;
    if (taps > 0) {
        const size_t per = (size_t)CinP * taps;
        const size_t co = i / per, r = i - co * per;
        const int t = (int)(r / CinP), c = (int)(r - (size_t)t * CinP);
        if (c >= Cin) return;
        o = (co * Cin + c) * taps + t;
    }
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += __ldg(part + (size_t)k * n + i);
    out[o] = s;
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint of reflect padding: gx[y][x] = sum over the padded positions that read x[y][x]
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_sources(int i, int n, int P, int* src) {      // padded indices that mirror onto i
    int k = 0;
    src[k++] = i + P;
    if (i >= 1 && i <= P) src[k++] = P - i;
    if (i <= n - 2 && i >= n - 1 - P) src[k++] = P + 2 * (n - 1) - i;
    return k;
}

__global__ void reflect_fold_kernel(const float* __restrict__ gxp, int gxp_ps, float* __restrict__ gx, int gx_ps, int B, int H, int W,
                                    int C, int P) {
    const size_t total = (size_t)B * H * W * C;
    const int Wp = W + 2 * P, Hp = H + 2 * P;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t p = i / C;
        const int x = (int)(p % W);
        p /= W;
        const int y = (int)(p % H), b = (int)(p / H);
        int sy[3], sx[3];
        const int ny = reflect_sources(y, H, P, sy), nx = reflect_sources(x, W, P, sx);
        float s = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int e = 0; e < nx; ++e) s += __ldg(gxp + (((size_t)b * Hp + sy[a]) * Wp + sx[e]) * gxp_ps + c);
        gx[(((size_t)b * H + y) * W + x) * gx_ps + c] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256, WG_TK = 16, WG_TN = 128;

struct WgradArgs {
    const float* x;      // [B,H,W,x_ps] NHWC
    const float* g;      // interior origin of the canvas: pixel (b,y,x) at ((b*g_img + y*g_row) + x) * g_ps
    float* part;         // [S][Cout][Ntot]
    int x_ps, g_ps;
    long long g_row, g_img;     // pixels
    int B, H, W, Cin, Cout, K, P, reflect;
    int CinP;            // im2col columns per tap: Cin, or Cin rounded up to 4 when the pixels of x can be read as float4
    int Ntot;            // K*K*CinP, n = tap*CinP + c  (columns with c >= Cin are dead)
    int x_vec;           // x pixels can be read as float4 channel quads
    int g_vec;           // g rows can be read as float4 (Cout, g_ps multiples of 4, 16-byte aligned base)
    long long npix, px_per_split;
};

__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// CTA tile TM x 128 (TM = 4*WM*MI output channels, 128 im2col columns), 16 pixels per staged chunk, 8 warps as WM x (8/WM).
// A warp covers 4 channel groups x 8 column groups, so a k-step costs it ONE shared-memory wavefront per 128-bit operand read
// (8 adjacent column quads = 128 B; 4 channel groups <= 128 B) for MI*NB FFMAs per thread: FFMA-bound, not shared-memory-bound.
// Thread (tm, tn): channels tm*MI .. +MI, column quads 4*tn + (128/NQ)*q, q < NQ = NB/4.
template <int MI, int NB, int WM>
__global__ void __launch_bounds__(WG_THREADS, (MI * NB <= 16 ? 4 : 2)) wgrad_kernel(const WgradArgs a) {
    constexpr int TM = 4 * WM * MI, WN = 8 / WM, TN = 8 * WN * NB, NQ = NB / 4, SB = TN / 16;
    static_assert(TN == WG_TN && TM * WG_TK / WG_THREADS >= 1, "tile shape");
    constexpr int SA = TM * WG_TK / WG_THREADS;          // A elements a thread stages per chunk
    __shared__ __align__(16) float As[WG_TK][TM];
    __shared__ __align__(16) float Bs[WG_TK][TN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tm = (warp % WM) * 4 + (lane >> 3), tn = (warp / WM) * 8 + (lane & 7);
    const int n0 = blockIdx.x * TN, m0 = blockIdx.y * TM;
    const long long p_begin = (long long)blockIdx.z * a.px_per_split;
    const long long p_end = min(p_begin + a.px_per_split, a.npix);

    // staging roles: pixel row krow = tid / 16 of the chunk; A -- channels ma .. ma+SA-1;  B -- column quads 4*(tid & 15) + 64*q
    const int krow = tid >> 4;
    const int ma = (tid & 15) * SA;
    int bcol[SB];                                         // c | (dy + 8) << 16 | (dx + 8) << 20, or -1 beyond the last column
#pragma unroll
    for (int j = 0; j < SB; ++j) {
        const int n = n0 + 64 * (j >> 2) + 4 * (tid & 15) + (j & 3);
        if (n < a.Ntot) {
            const int t = n / a.CinP;
            bcol[j] = (n - t * a.CinP) | ((t / a.K - a.P + 8) << 16) | ((t % a.K - a.P + 8) << 20);
        } else {
            bcol[j] = -1;
        }
    }
    float ra[SA], rb[SB];
    auto fetch = [&](long long pbase) {
        const long long p = pbase + krow;
        const bool live = p < p_end;
        int b = 0, y = 0, x = 0;
        if (live) {
            const unsigned hw = (unsigned)(a.H * a.W);
            b = (int)(p / hw);
            const unsigned r = (unsigned)(p - (long long)b * hw);
            y = (int)(r / (unsigned)a.W);
            x = (int)(r - (unsigned)y * (unsigned)a.W);
        }
        const float* gp = a.g + ((size_t)b * a.g_img + (size_t)y * a.g_row + x) * a.g_ps + m0 + ma;
        if (SA >= 4 && a.g_vec && m0 + ma + SA <= a.Cout) {
#pragma unroll
            for (int i = 0; i + 3 < SA; i += 4) {
                const float4 v = live ? __ldg(reinterpret_cast<const float4*>(gp + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                ra[i] = v.x; ra[i + 1] = v.y; ra[i + 2] = v.z; ra[i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < SA; ++i) ra[i] = (live && m0 + ma + i < a.Cout) ? __ldg(gp + i) : 0.f;
        }
        const float* xb = a.x + (size_t)b * a.H * a.W * a.x_ps;
        if (a.x_vec) {                                   // a column quad = 4 adjacent channels of one tap: one 128-bit load
#pragma unroll
            for (int j = 0; j < SB; j += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && bcol[j] >= 0) {
                    int yy = y + ((bcol[j] >> 16) & 15) - 8, xx = x + ((bcol[j] >> 20) & 15) - 8;
                    bool ok = true;
                    if (a.reflect) {
                        yy = reflect_index(yy, a.H);
                        xx = reflect_index(xx, a.W);
                    } else {
                        ok = (unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W;
                    }
                    if (ok) v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(yy * a.W + xx) * a.x_ps + (bcol[j] & 0xffff)));
                }
                rb[j] = v.x; rb[j + 1] = v.y; rb[j + 2] = v.z; rb[j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                float v = 0.f;
                if (live && bcol[j] >= 0) {
                    int yy = y + ((bcol[j] >> 16) & 15) - 8, xx = x + ((bcol[j] >> 20) & 15) - 8;
                    bool ok = true;
                    if (a.reflect) {
                        yy = reflect_index(yy, a.H);
                        xx = reflect_index(xx, a.W);
                    } else {
                        ok = (unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W;
                    }
                    if (ok) v = __ldg(xb + (size_t)(yy * a.W + xx) * a.x_ps + (bcol[j] & 0xffff));
                }
                rb[j] = v;
            }
        }
    };

    float acc[MI][NB];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j] = 0.f;

    if (p_begin < p_end) fetch(p_begin);
    for (long long pb = p_begin; pb < p_end; pb += WG_TK) {
#pragma unroll
        for (int i = 0; i < SA; ++i) As[krow][ma + i] = ra[i];
#pragma unroll
        for (int q = 0; q < SB / 4; ++q)
            *reinterpret_cast<float4*>(&Bs[krow][64 * q + 4 * (tid & 15)]) = make_float4(rb[4 * q], rb[4 * q + 1], rb[4 * q + 2], rb[4 * q + 3]);
        __syncthreads();
        if (pb + WG_TK < p_end) fetch(pb + WG_TK);
#pragma unroll
        for (int k = 0; k < WG_TK; ++k) {
            float av[MI];
#pragma unroll
            for (int i = 0; i < MI; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As[k][tm * MI + i]);
                av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][(TN / NQ) * q + 4 * tn]);
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    acc[i][4 * q + 0] = fmaf(av[i], bv.x, acc[i][4 * q + 0]);
                    acc[i][4 * q + 1] = fmaf(av[i], bv.y, acc[i][4 * q + 1]);
                    acc[i][4 * q + 2] = fmaf(av[i], bv.z, acc[i][4 * q + 2]);
                    acc[i][4 * q + 3] = fmaf(av[i], bv.w, acc[i][4 * q + 3]);
                }
            }
        }
        __syncthreads();
    }
    float* out = a.part + (size_t)blockIdx.z * a.Cout * a.Ntot;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + tm * MI + i;
        if (m >= a.Cout) continue;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int n = n0 + (TN / NQ) * (j >> 2) + 4 * tn + (j & 3);
            if (n < a.Ntot) out[(size_t)m * a.Ntot + n] = acc[i][j];
        }
    }
}

static int wgrad_tile_m(int Cout) { return Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128); }

static int wgrad_splits(int Cout, int Ntot, long long npix) {
    const long long tiles = (long long)ceil_div(Ntot, WG_TN) * ceil_div(Cout, wgrad_tile_m(Cout));
    const long long want = 4LL * std::max(sm_count(), 1);                 // ~4 CTAs per SM in flight
    long long S = std::max(1LL, (want + tiles - 1) / tiles);
    const long long max_s = std::max(1LL, npix / (8 * WG_TK));            // at least 8 chunks per split
    return (int)std::min(S, max_s);
}

// ------------------------------------------------------------------------------------------------------------------
// the other differentiable steps of FusionNet's training forward (src/fusion_net/fusion_net.py:52-77)
// ------------------------------------------------------------------------------------------------------------------
// nn.MaxPool2d(2, 2) backward: the gradient of a window goes to its FIRST maximum in scan order (ATen's rule: `val > max`)
__global__ void __launch_bounds__(256) max_pool2_bwd_kernel(const float* __restrict__ x, int x_ps, const float* __restrict__ gy, int gy_ps,
                                                           float* __restrict__ gx, int gx_ps, int B, int Hi, int Wi, int C) {
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const size_t total = (size_t)B * Hi * Wi * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t p = i / C;
        const int w = (int)(p % Wi);
        p /= Wi;
        const int h = (int)(p % Hi), b = (int)(p / Hi);
        float g = 0.f;
        const int oy = h >> 1, ox = w >> 1;
        if (oy < Ho && ox < Wo) {
            const float* win = x + (((size_t)b * Hi + 2 * oy) * Wi + 2 * ox) * x_ps + c;
            float m = __ldg(win);
            int arg = 0;
            const float v1 = __ldg(win + x_ps), v2 = __ldg(win + (size_t)Wi * x_ps), v3 = __ldg(win + (size_t)(Wi + 1) * x_ps);
            if (v1 > m) { m = v1; arg = 1; }
            if (v2 > m) { m = v2; arg = 2; }
            if (v3 > m) { m = v3; arg = 3; }
            if (arg == ((h & 1) << 1 | (w & 1))) g = __ldg(gy + (((size_t)b * Ho + oy) * Wo + ox) * gy_ps + c);
        }
        gx[(((size_t)b * Hi + h) * Wi + w) * gx_ps + c] = g;
    }
}

// nn.AvgPool2d(2, 2) backward: every pixel of a window receives gy / 4; an odd last row / column gets zero
__global__ void __launch_bounds__(256) avg_pool2_bwd_kernel(const float* __restrict__ gy, int gy_ps, float* __restrict__ gx, int gx_ps,
                                                           int B, int Hi, int Wi, int C) {
    const int Ho = Hi >> 1, Wo = Wi >> 1;
    const size_t total = (size_t)B * Hi * Wi * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t p = i / C;
        const int w = (int)(p % Wi);
        p /= Wi;
        const int h = (int)(p % Hi), b = (int)(p / Hi);
        const int oy = h >> 1, ox = w >> 1;
        const float g = (oy < Ho && ox < Wo) ? 0.25f * __ldg(gy + (((size_t)b * Ho + oy) * Wo + ox) * gy_ps + c) : 0.f;
        gx[(((size_t)b * Hi + h) * Wi + w) * gx_ps + c] = g;
    }
}

// adjoint of fvfi_resize_bilinear_nhwc_fused for upsampling factors <= 3 per axis, in gather form: input pixel (i, j) collects
// w_y(o, i) * w_x(p, j) * gy[o, p] over the outputs whose two source rows / columns include it (weights from the SAME
// bilinear_src as the forward), times relu'(x) when the forward resampled max(x, 0).
constexpr int RB_MAXC = 10;        // candidate outputs per axis: 2 * 3 + margins
__device__ __forceinline__ int resize_adjoint_weights(int i, int n_in, int n_out, float scale, int align, int* o0, float* w) {
    // outputs o with source coordinate in (i - 1, i + 1): o in [(i - 1) / scale - 1, (i + 1) / scale + 1]
    const float inv = scale > 0.f ? 1.f / scale : 0.f;
    int lo = max(0, (int)floorf(((float)i - 1.f) * inv) - 1);
    int hi = min(n_out - 1, (int)ceilf(((float)i + 1.f) * inv) + 1);
    if (scale == 0.f) { lo = 0; hi = n_out - 1; }
    if (hi - lo + 1 > RB_MAXC) hi = lo + RB_MAXC - 1;
    *o0 = lo;
    const int n = hi - lo + 1;
    for (int k = 0; k < RB_MAXC; ++k) {
        float ww = 0.f;
        if (k < n) {
            int i0, i1;
            float l;
            bilinear_src(lo + k, scale, align, n_in, i0, i1, l);
            if (i0 == i) ww += 1.f - l;
            if (i1 == i) ww += l;
        }
        w[k] = ww;
    }
    return n;
}

__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const float* __restrict__ gy, int gy_ps, const float* __restrict__ x,
                                                                 int x_ps, float* __restrict__ gx, int gx_ps, int B, int Hi, int Wi, int Ho,
                                                                 int Wo, int C, int align, float sy, float sx, int relu_input) {
    // one thread per (input pixel, channel quad-or-single); the block's threads share nothing, the weights are per pixel
    const int cq = (C + 3) / 4;
    const size_t total = (size_t)B * Hi * Wi * cq;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(q % cq) * 4;
        size_t p = q / cq;
        const int j = (int)(p % Wi);
        p /= Wi;
        const int i = (int)(p % Hi), b = (int)(p / Hi);
        float wy[RB_MAXC], wx[RB_MAXC];
        int oy0, ox0;
        const int ny = resize_adjoint_weights(i, Hi, Ho, sy, align, &oy0, wy);
        const int nx = resize_adjoint_weights(j, Wi, Wo, sx, align, &ox0, wx);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int nc = min(4, C - c0);
        for (int a = 0; a < ny; ++a) {
            if (wy[a] == 0.f) continue;
            const float* row = gy + (((size_t)b * Ho + oy0 + a) * Wo + ox0) * gy_ps + c0;
            for (int e = 0; e < nx; ++e) {
                const float w = wy[a] * wx[e];
                if (w == 0.f) continue;
                for (int c = 0; c < nc; ++c) acc[c] = fmaf(w, __ldg(row + (size_t)e * gy_ps + c), acc[c]);
            }
        }
        const size_t pix = ((size_t)b * Hi + i) * Wi + j;
        for (int c = 0; c < nc; ++c) {
            float v = acc[c];
            if (relu_input && !(__ldg(x + pix * x_ps + c0 + c) > 0.f)) v = 0.f;
            gx[pix * gx_ps + c0 + c] = v;
        }
    }
}

// out = clamp(base + tanh(x), 0, 1) (fvfi_fusion_blend; fusion_net.py:67-77):  gx = gout * [0 < base + tanh x < 1] * (1 - tanh^2 x),
// gbase (optional) = gout * [..]
__global__ void __launch_bounds__(256) fusion_blend_bwd_kernel(const float* __restrict__ base, const float* __restrict__ x,
                                                              const float* __restrict__ gout, float* __restrict__ gx,
                                                              float* __restrict__ gbase, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float t = tanhf(__ldg(x + i));
        const float o = __ldg(base + i) + t;
        const float g = (o > 0.f && o < 1.f) ? __ldg(gout + i) : 0.f;
        gx[i] = g * (1.f - t * t);
        if (gbase) gbase[i] = g;
    }
}


static int grad_act_blocks(int rows) { return std::min(rows, 8 * std::max(sm_count(), 1)); }

}  // namespace fvfi

using namespace fvfi;

extern "C" {

size_t fvfi_conv2d_grad_act_workspace_floats(int B, int H, int W, int C, int border) {
    (void)W;
    return (size_t)grad_act_blocks(B * (H + 2 * border)) * (size_t)C;
}

int fvfi_conv2d_grad_act(const float* gy, int gy_pixel_stride, const float* y, int y_pixel_stride, float* g_canvas, int B, int H,
                         int W, int C, int border, int activation, float* gbias, float* workspace, void* stream) {
    FVFI_CHECK_ARG(gy && g_canvas && B > 0 && H > 0 && W > 0 && C > 0 && border >= 0, "fvfi_conv2d_grad_act: bad arguments");
    FVFI_CHECK_ARG(activation == BACT_NONE || (y != nullptr && activation >= 0 && activation <= BACT_SIGMOID),
                   "fvfi_conv2d_grad_act: activation %d needs the saved output y (0 none, 1 ReLU, 2 ELU, 3 tanh, 4 sigmoid)", activation);
    FVFI_CHECK_ARG(gy_pixel_stride >= C && (y == nullptr || y_pixel_stride >= C), "fvfi_conv2d_grad_act: pixel stride < C");
    FVFI_CHECK_ARG(gbias == nullptr || workspace != nullptr, "fvfi_conv2d_grad_act: the bias gradient needs the workspace");
    const bool vec = C % 4 == 0 && gy_pixel_stride % 4 == 0 && (y == nullptr || y_pixel_stride % 4 == 0) &&
                     (((size_t)gy | (size_t)g_canvas | (size_t)y) & 15) == 0;
    const int items = vec ? C / 4 : C;
    int CX = 8;
    while (CX < items && CX < GA_THREADS) CX *= 2;
    FVFI_CHECK_ARG(CX * GA_MAX_CPT >= items, "fvfi_conv2d_grad_act: C = %d too large", C);
    const int PY = GA_THREADS / CX;
    const int blocks = grad_act_blocks(B * (H + 2 * border));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = gbias ? (size_t)PY * C * sizeof(float) : 0;
    if (vec)
        grad_act_kernel<true><<<blocks, dim3(CX, PY), smem, st>>>(gy, gy_pixel_stride, y, y_pixel_stride, g_canvas, B, H, W, C, border,
                                                                 activation, gbias ? workspace : nullptr);
    else
        grad_act_kernel<false><<<blocks, dim3(CX, PY), smem, st>>>(gy, gy_pixel_stride, y, y_pixel_stride, g_canvas, B, H, W, C, border,
                                                                  activation, gbias ? workspace : nullptr);
    FVFI_LAUNCH_CHECK();
    if (gbias) {
        sum_partials_kernel<<<ceil_div(C, 128), 128, 0, st>>>(workspace, blocks, (size_t)C, gbias, 0, 0, 0);
        FVFI_LAUNCH_CHECK();
    }
    return FVFI_OK;
}

int fvfi_reflect_pad_backward_nhwc(const float* gxp, int gxp_pixel_stride, float* gx, int gx_pixel_stride, int B, int H, int W, int C,
                                   int P, void* stream) {
    FVFI_CHECK_ARG(gxp && gx && B > 0 && C > 0 && P >= 0 && H > P && W > P, "fvfi_reflect_pad_backward_nhwc: bad arguments (needs H, W > P)");
    FVFI_CHECK_ARG(gxp_pixel_stride >= C && gx_pixel_stride >= C, "fvfi_reflect_pad_backward_nhwc: pixel stride < C");
    const size_t total = (size_t)B * H * W * C;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    reflect_fold_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gxp, gxp_pixel_stride, gx, gx_pixel_stride, B, H, W, C, P);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

size_t fvfi_conv2d_wgrad_workspace_floats(int B, int H, int W, int Cin, int Cout, int K) {
    // columns per tap: Cin, or Cin rounded up to 4 when the call finds x readable as float4 -- cover both
    const long long npix = (long long)B * H * W;
    const int n1 = K * K * Cin, n2 = K * K * ((Cin + 3) / 4 * 4);
    return std::max((size_t)wgrad_splits(Cout, n1, npix) * (size_t)n1, (size_t)wgrad_splits(Cout, n2, npix) * (size_t)n2) * (size_t)Cout;
}

int fvfi_conv2d_wgrad_nhwc(const float* x, int x_pixel_stride, const float* g, int g_pixel_stride, long long g_row_pixels,
                           long long g_image_pixels, float* gw_oihw, int B, int H, int W, int Cin, int Cout, int K, int pad_mode,
                           float* workspace, void* stream) {
    FVFI_CHECK_ARG(x && g && gw_oihw && workspace && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "fvfi_conv2d_wgrad_nhwc: bad arguments");
    FVFI_CHECK_ARG(K >= 1 && (K & 1) && (pad_mode == 0 || pad_mode == 1), "fvfi_conv2d_wgrad_nhwc: odd K, pad_mode 0 zeros / 1 reflect");
    FVFI_CHECK_ARG(pad_mode == 0 || (H > K / 2 && W > K / 2), "fvfi_conv2d_wgrad_nhwc: reflect padding needs H, W > K/2");
    FVFI_CHECK_ARG(x_pixel_stride >= Cin && g_pixel_stride >= Cout && g_row_pixels >= W && g_image_pixels >= (long long)H * g_row_pixels - (g_row_pixels - W),
                   "fvfi_conv2d_wgrad_nhwc: strides too small");
    WgradArgs a;
    a.x = x; a.g = g; a.part = workspace;
    a.x_ps = x_pixel_stride; a.g_ps = g_pixel_stride; a.g_row = g_row_pixels; a.g_img = g_image_pixels;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.K = K; a.P = K / 2; a.reflect = pad_mode;
    a.x_vec = (x_pixel_stride % 4 == 0 && x_pixel_stride >= (Cin + 3) / 4 * 4 && ((size_t)x & 15) == 0) ? 1 : 0;
    a.CinP = a.x_vec ? (Cin + 3) / 4 * 4 : Cin;
    a.Ntot = K * K * a.CinP;
    a.npix = (long long)B * H * W;
    const int S = wgrad_splits(Cout, a.Ntot, a.npix);
    long long per = (a.npix + S - 1) / S;
    per = (per + WG_TK - 1) / WG_TK * WG_TK;
    a.px_per_split = per;
    a.g_vec = (Cout % 4 == 0 && g_pixel_stride % 4 == 0 && ((size_t)g & 15) == 0) ? 1 : 0;
    const int tm = wgrad_tile_m(Cout);
    const dim3 grid(ceil_div(a.Ntot, WG_TN), ceil_div(Cout, tm), S);
    cudaStream_t st = (cudaStream_t)stream;
    if (tm == 32) wgrad_kernel<4, 4, 2><<<grid, WG_THREADS, 0, st>>>(a);
    else if (tm == 64) wgrad_kernel<4, 8, 4><<<grid, WG_THREADS, 0, st>>>(a);
    else wgrad_kernel<8, 8, 4><<<grid, WG_THREADS, 0, st>>>(a);
    FVFI_LAUNCH_CHECK();
    const size_t n = (size_t)Cout * a.Ntot;
    sum_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, S, n, gw_oihw, Cin, a.CinP, K * K);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

int fvfi_max_pool2_backward_nhwc(const float* x, int x_pixel_stride, const float* gy, int gy_pixel_stride, float* gx, int gx_pixel_stride,
                                 int B, int Hi, int Wi, int C, void* stream) {
    FVFI_CHECK_ARG(x && gy && gx && B > 0 && Hi > 1 && Wi > 1 && C > 0, "fvfi_max_pool2_backward_nhwc: bad arguments");
    FVFI_CHECK_ARG(x_pixel_stride >= C && gy_pixel_stride >= C && gx_pixel_stride >= C, "fvfi_max_pool2_backward_nhwc: pixel stride < C");
    const size_t total = (size_t)B * Hi * Wi * C;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    max_pool2_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_pixel_stride, gy, gy_pixel_stride, gx, gx_pixel_stride, B, Hi, Wi, C);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

int fvfi_avg_pool2_backward_nhwc(const float* gy, int gy_pixel_stride, float* gx, int gx_pixel_stride, int B, int Hi, int Wi, int C,
                                 void* stream) {
    FVFI_CHECK_ARG(gy && gx && B > 0 && Hi > 1 && Wi > 1 && C > 0, "fvfi_avg_pool2_backward_nhwc: bad arguments");
    FVFI_CHECK_ARG(gy_pixel_stride >= C && gx_pixel_stride >= C, "fvfi_avg_pool2_backward_nhwc: pixel stride < C");
    const size_t total = (size_t)B * Hi * Wi * C;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    avg_pool2_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gy, gy_pixel_stride, gx, gx_pixel_stride, B, Hi, Wi, C);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

int fvfi_resize_bilinear_backward_nhwc(const float* gy, int gy_pixel_stride, const float* x, int x_pixel_stride, float* gx,
                                       int gx_pixel_stride, int B, int Hi, int Wi, int Ho, int Wo, int C, int align_corners,
                                       int relu_input, void* stream) {
    FVFI_CHECK_ARG(gy && gx && B > 0 && Hi > 0 && Wi > 0 && C > 0, "fvfi_resize_bilinear_backward_nhwc: bad arguments");
    FVFI_CHECK_ARG(Ho >= Hi && Wo >= Wi && Ho <= 3 * Hi && Wo <= 3 * Wi, "fvfi_resize_bilinear_backward_nhwc: upsampling by at most 3 per axis");
    FVFI_CHECK_ARG(!relu_input || x != nullptr, "fvfi_resize_bilinear_backward_nhwc: relu_input needs the forward input x");
    FVFI_CHECK_ARG(gy_pixel_stride >= C && gx_pixel_stride >= C && (!x || x_pixel_stride >= C), "fvfi_resize_bilinear_backward_nhwc: pixel stride < C");
    const size_t total = (size_t)B * Hi * Wi * ((C + 3) / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    resize_bilinear_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gy, gy_pixel_stride, x, x_pixel_stride, gx, gx_pixel_stride, B, Hi, Wi,
                                                                        Ho, Wo, C, align_corners, bilinear_scale(Hi, Ho, align_corners),
                                                                        bilinear_scale(Wi, Wo, align_corners), relu_input);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

int fvfi_fusion_blend_backward(const float* base, const float* x, const float* gout, float* gx, float* gbase, size_t n, void* stream) {
    FVFI_CHECK_ARG(base && x && gout && gx && n > 0, "fvfi_fusion_blend_backward: bad arguments");
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
    fusion_blend_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(base, x, gout, gx, gbase, n);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

}  // extern "C"
