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
constexpr int GA_MAX_CPT = 8;       // channels per thread (C <= 8 * 32 at CX = 32; CX grows with C)

__global__ void __launch_bounds__(GA_THREADS) grad_act_kernel(const float* __restrict__ gy, int gy_ps, const float* __restrict__ y,
                                                             int y_ps, float* __restrict__ canvas, int B, int H, int W, int C,
                                                             int border, int act, float* __restrict__ bias_part,
                                                             long long px_per_block) {
    extern __shared__ float red[];                       // [PY][C]
    const int CX = blockDim.x, PY = blockDim.y;
    const int Hc = H + 2 * border, Wc = W + 2 * border;
    const long long npix = (long long)B * Hc * Wc;
    const long long p0 = (long long)blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, npix);
    float sum[GA_MAX_CPT];
#pragma unroll
    for (int i = 0; i < GA_MAX_CPT; ++i) sum[i] = 0.f;
    for (long long p = p0 + threadIdx.y; p < p1; p += PY) {
        const int b = (int)(p / ((long long)Hc * Wc));
        const int r = (int)(p - (long long)b * Hc * Wc);
        const int yy = r / Wc - border, xx = r % Wc - border;
        const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const size_t src = ((size_t)b * H + (inside ? yy : 0)) * W + (inside ? xx : 0);
#pragma unroll
        for (int i = 0; i < GA_MAX_CPT; ++i) {
            const int c = threadIdx.x + i * CX;
            if (c < C) {
                float v = 0.f;
                if (inside) v = __ldg(gy + src * gy_ps + c) * (act ? act_grad_from_output(__ldg(y + src * y_ps + c), act) : 1.f);
                canvas[(size_t)p * C + c] = v;
                sum[i] += v;
            }
        }
    }
    if (bias_part == nullptr) return;
#pragma unroll
    for (int i = 0; i < GA_MAX_CPT; ++i) {
        const int c = threadIdx.x + i * CX;
        if (c < C) red[threadIdx.y * C + c] = sum[i];
    }
    __syncthreads();
    for (int c = threadIdx.y * CX + threadIdx.x; c < C; c += CX * PY) {
        float s = 0.f;
        for (int j = 0; j < PY; ++j) s += red[j * C + c];        // fixed order
        bias_part[(size_t)blockIdx.x * C + c] = s;
    }
}

// out[i] = sum_s part[s * n + i], fixed order; remap: wgrad partials are [Cout][tap][Cin], the gradient is OIHW [Cout][Cin][tap]
__global__ void sum_partials_kernel(const float* __restrict__ part, int S, size_t n, float* __restrict__ out, int Cin, int taps) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += __ldg(part + (size_t)k * n + i);
    size_t o = i;
    if (taps > 0) {
        const size_t per = (size_t)Cin * taps;
        const size_t co = i / per, r = i - co * per;
        const int t = (int)(r / Cin), c = (int)(r - (size_t)t * Cin);
        o = co * per + (size_t)c * taps + t;
    }
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
constexpr int WG_THREADS = 256, WG_TN = 64, WG_TK = 16;

struct WgradArgs {
    const float* x;      // [B,H,W,x_ps] NHWC
    const float* g;      // interior origin of the canvas: pixel (b,y,x) at ((b*g_img + y*g_row) + x) * g_ps
    float* part;         // [S][Cout][Ntot]
    int x_ps, g_ps;
    long long g_row, g_img;     // pixels
    int B, H, W, Cin, Cout, K, P, reflect;
    int Ntot;            // K*K*Cin, n = tap*Cin + c
    long long npix, px_per_split;
};

__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

template <int MI>      // output channels per thread; the CTA tile is (16*MI) x 64
__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const WgradArgs a) {
    constexpr int TM = 16 * MI;
    constexpr int A_PER = TM * WG_TK / WG_THREADS;          // A elements a thread stages per chunk (MI)
    __shared__ __align__(16) float As[WG_TK][TM];
    __shared__ __align__(16) float Bs[WG_TK][WG_TN];
    const int tid = threadIdx.x;
    const int tn = tid & 15, tm = tid >> 4;
    const int n0 = blockIdx.x * WG_TN, m0 = blockIdx.y * TM;
    const long long p_begin = (long long)blockIdx.z * a.px_per_split;
    const long long p_end = min(p_begin + a.px_per_split, a.npix);

    // staging roles: A -- row ka = tid / 16, channels ma .. ma+MI-1;   B -- row kb = tid / 16, columns nb .. nb+3
    const int krow = tid >> 4;
    const int ma = (tid & 15) * A_PER;
    const int nb = (tid & 15) * 4;
    int bdy[4], bdx[4], bc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + nb + j;
        if (n < a.Ntot) {
            const int t = n / a.Cin;
            bc[j] = n - t * a.Cin;
            bdy[j] = t / a.K - a.P;
            bdx[j] = t % a.K - a.P;
        } else {
            bc[j] = -1; bdy[j] = 0; bdx[j] = 0;
        }
    }
    float ra[A_PER], rb[4];
    auto fetch = [&](long long pbase) {
        const long long p = pbase + krow;
        const bool live = p < p_end;
        int b = 0, y = 0, x = 0;
        if (live) {
            const long long hw = (long long)a.H * a.W;
            b = (int)(p / hw);
            const int r = (int)(p - (long long)b * hw);
            y = r / a.W;
            x = r - y * a.W;
        }
        const float* gp = a.g + ((size_t)b * a.g_img + (size_t)y * a.g_row + x) * a.g_ps + m0 + ma;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) ra[i] = (live && m0 + ma + i < a.Cout) ? __ldg(gp + i) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = 0.f;
            if (live && bc[j] >= 0) {
                int yy = y + bdy[j], xx = x + bdx[j];
                bool ok = true;
                if (a.reflect) {
                    yy = reflect_index(yy, a.H);
                    xx = reflect_index(xx, a.W);
                } else {
                    ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
                }
                if (ok) v = __ldg(a.x + (((size_t)b * a.H + yy) * a.W + xx) * a.x_ps + bc[j]);
            }
            rb[j] = v;
        }
    };

    float acc[MI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (p_begin < p_end) fetch(p_begin);
    for (long long pb = p_begin; pb < p_end; pb += WG_TK) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) As[krow][ma + i] = ra[i];
        *reinterpret_cast<float4*>(&Bs[krow][nb]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        __syncthreads();
        if (pb + WG_TK < p_end) fetch(pb + WG_TK);
#pragma unroll
        for (int k = 0; k < WG_TK; ++k) {
            float av[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) av[i] = As[k][tm * MI + i];
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                acc[i][0] = fmaf(av[i], bv.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], bv.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], bv.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], bv.w, acc[i][3]);
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
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n < a.Ntot) out[(size_t)m * a.Ntot + n] = acc[i][j];
        }
    }
}

static int wgrad_tile_m(int Cout) { return Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128); }

static int wgrad_splits(int Cout, int Ntot, long long npix) {
    const int tm = wgrad_tile_m(Cout);
    const long long tiles = (long long)ceil_div(Ntot, WG_TN) * ceil_div(Cout, tm);
    const long long want = 4LL * std::max(sm_count(), 1);                 // ~4 CTAs per SM in flight
    long long S = std::max(1LL, (want + tiles - 1) / tiles);
    const long long max_s = std::max(1LL, npix / (8 * WG_TK));            // at least 8 chunks per split
    return (int)std::min(S, max_s);
}

static int grad_act_blocks(long long npix) { return (int)std::min<long long>(std::max<long long>(npix / 512, 1), 1184); }

}  // namespace fvfi

using namespace fvfi;

extern "C" {

size_t fvfi_conv2d_grad_act_workspace_floats(int B, int H, int W, int C, int border) {
    const long long npix = (long long)B * (H + 2 * border) * (W + 2 * border);
    return (size_t)grad_act_blocks(npix) * (size_t)C;
}

int fvfi_conv2d_grad_act(const float* gy, int gy_pixel_stride, const float* y, int y_pixel_stride, float* g_canvas, int B, int H,
                         int W, int C, int border, int activation, float* gbias, float* workspace, void* stream) {
    FVFI_CHECK_ARG(gy && g_canvas && B > 0 && H > 0 && W > 0 && C > 0 && border >= 0, "fvfi_conv2d_grad_act: bad arguments");
    FVFI_CHECK_ARG(activation == BACT_NONE || (y != nullptr && activation >= 0 && activation <= BACT_SIGMOID),
                   "fvfi_conv2d_grad_act: activation %d needs the saved output y (0 none, 1 ReLU, 2 ELU, 3 tanh, 4 sigmoid)", activation);
    FVFI_CHECK_ARG(gy_pixel_stride >= C && (y == nullptr || y_pixel_stride >= C), "fvfi_conv2d_grad_act: pixel stride < C");
    FVFI_CHECK_ARG(gbias == nullptr || workspace != nullptr, "fvfi_conv2d_grad_act: the bias gradient needs the workspace");
    int CX = 32;
    while (CX * GA_MAX_CPT < C && CX < GA_THREADS) CX *= 2;
    FVFI_CHECK_ARG(CX * GA_MAX_CPT >= C, "fvfi_conv2d_grad_act: C = %d > %d", C, GA_THREADS * GA_MAX_CPT);
    const int PY = GA_THREADS / CX;
    const long long npix = (long long)B * (H + 2 * border) * (W + 2 * border);
    const int blocks = grad_act_blocks(npix);
    const long long per = (npix + blocks - 1) / blocks;
    cudaStream_t st = (cudaStream_t)stream;
    grad_act_kernel<<<blocks, dim3(CX, PY), gbias ? (size_t)PY * C * sizeof(float) : 0, st>>>(
        gy, gy_pixel_stride, y, y_pixel_stride, g_canvas, B, H, W, C, border, activation, gbias ? workspace : nullptr, per);
    FVFI_LAUNCH_CHECK();
    if (gbias) {
        sum_partials_kernel<<<ceil_div(C, 128), 128, 0, st>>>(workspace, blocks, (size_t)C, gbias, 0, 0);
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
    const int Ntot = K * K * Cin;
    return (size_t)wgrad_splits(Cout, Ntot, (long long)B * H * W) * (size_t)Cout * (size_t)Ntot;
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
    a.Ntot = K * K * Cin;
    a.npix = (long long)B * H * W;
    const int S = wgrad_splits(Cout, a.Ntot, a.npix);
    long long per = (a.npix + S - 1) / S;
    per = (per + WG_TK - 1) / WG_TK * WG_TK;
    a.px_per_split = per;
    const int tm = wgrad_tile_m(Cout);
    const dim3 grid(ceil_div(a.Ntot, WG_TN), ceil_div(Cout, tm), S);
    cudaStream_t st = (cudaStream_t)stream;
    if (tm == 32) wgrad_kernel<2><<<grid, WG_THREADS, 0, st>>>(a);
    else if (tm == 64) wgrad_kernel<4><<<grid, WG_THREADS, 0, st>>>(a);
    else wgrad_kernel<8><<<grid, WG_THREADS, 0, st>>>(a);
    FVFI_LAUNCH_CHECK();
    const size_t n = (size_t)Cout * a.Ntot;
    sum_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, S, n, gw_oihw, Cin, K * K);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

}  // extern "C"
