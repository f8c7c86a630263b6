// conv_direct.cu -- the two convolution shapes of the path that are pure HBM streams, on CUDA cores (fp32 FFMA):
//
//   fvfi_conv1x1_nhwc            1x1 convolution with Cout <= 8: PhaseNet's per-level prediction 64 -> 8 + tanh
//                                (src/phase_net/phase_net.py:197-200) and FusionNet's last 32 -> 3 (src/fusion_net/fusion_net.py:36).
//                                512 FMA per 256 B pixel: the tensor-core kernel spends a 128 x 16 MMA tile per 8 useful
//                                columns and runs these layers at ~2 TB/s; here one thread owns a pixel (256-bit loads,
//                                weights broadcast from shared memory) and the layer is one read of the activation.
//   fvfi_upsample2_tapsum        the occlusion head's tail  Upsample(x2, bilinear, align_corners=True) -> Conv2d(64, 1, 3, pad 1)
//                                -> Sigmoid  (src/fusion_net/fusion_adacofnet.py:50-59 with last_in = 64).  Both steps are
//                                linear, so the 64 channels are contracted FIRST, at half resolution: z_t = sum_c w[c,t] x_c
//                                for the nine taps t (a 64 -> 9 1x1 convolution, tensor-core kernel), and this kernel evaluates
//                                out(i,j) = act(b + sum_t [p_t inside] bilerp(z_t, p_t)),  p_t = (i + dy_t - 1, j + dx_t - 1):
//                                the 64-channel full-resolution tensor (4.3 GB written + read at 1080p, batch 8) never exists.
#include <algorithm>

#include "common.cuh"

namespace fvfi {

enum { DACT_NONE = 0, DACT_RELU = 1, DACT_ELU = 2, DACT_TANH = 3, DACT_SIGMOID = 4 };

// same forms as the tensor-core epilogue (conv_tc.cu): hardware ex2, absolute error <= 1e-6
__device__ __forceinline__ float dact(float v, int act) {
    switch (act) {
        case DACT_RELU: return fmaxf(v, 0.f);
        case DACT_ELU: return v > 0.f ? v : __expf(v) - 1.f;
        case DACT_TANH: {
            const float t = __expf(2.f * fminf(fmaxf(v, -15.f), 15.f));
            return __fdividef(t - 1.f, t + 1.f);
        }
        case DACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-fmaxf(v, -80.f)));
        default: return v;
    }
}

__device__ __forceinline__ void ldg256s(const float* p, float* v) {      // 256-bit streaming load: one 32 B sector per lane
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

constexpr int C1_THREADS = 256;
constexpr int C1_PX = 2;            // pixels per thread and iteration (weights are read once for both)

// x [npix, ldx] (Cin used), w_s [Cin][8] in shared memory (zero beyond Cout), y [npix, ldy] (Cout written)
template <int CIN8>                 // Cin / 8, compile-time so the channel loop unrolls around the 256-bit loads
__global__ void __launch_bounds__(C1_THREADS) conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   size_t npix, int Cout, int ldx, int ldy, int act) {
    __shared__ float4 w_s[CIN8 * 8 * 2];
    __shared__ float b_s[8];
    constexpr int CIN = CIN8 * 8;
    for (int i = threadIdx.x; i < CIN * 8; i += C1_THREADS) {
        const int c = i >> 3, o = i & 7;
        ((float*)w_s)[i] = o < Cout ? __ldg(w + (size_t)o * CIN + c) : 0.f;
    }
    if (threadIdx.x < 8) b_s[threadIdx.x] = (bias && (int)threadIdx.x < Cout) ? __ldg(bias + threadIdx.x) : 0.f;
    __syncthreads();
    const size_t T = (size_t)gridDim.x * C1_THREADS;
    const bool vec_out = (Cout == 8) && ((ldy & 7) == 0) && ((((size_t)y) & 31) == 0);
    for (size_t p0 = (size_t)blockIdx.x * C1_THREADS + threadIdx.x; p0 < npix; p0 += C1_PX * T) {
        float acc[C1_PX][8];
#pragma unroll
        for (int u = 0; u < C1_PX; ++u)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[u][o] = b_s[o];
#pragma unroll
        for (int k = 0; k < CIN8; ++k) {
            float v[C1_PX][8];
#pragma unroll
            for (int u = 0; u < C1_PX; ++u) {
                const size_t p = p0 + (size_t)u * T;
                if (p < npix) ldg256s(x + p * (size_t)ldx + k * 8, v[u]);
                else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 wa = w_s[(k * 8 + e) * 2], wb = w_s[(k * 8 + e) * 2 + 1];      // broadcast: one wavefront each
#pragma unroll
                for (int u = 0; u < C1_PX; ++u) {
                    const float xv = v[u][e];
                    acc[u][0] = fmaf(xv, wa.x, acc[u][0]); acc[u][1] = fmaf(xv, wa.y, acc[u][1]);
                    acc[u][2] = fmaf(xv, wa.z, acc[u][2]); acc[u][3] = fmaf(xv, wa.w, acc[u][3]);
                    acc[u][4] = fmaf(xv, wb.x, acc[u][4]); acc[u][5] = fmaf(xv, wb.y, acc[u][5]);
                    acc[u][6] = fmaf(xv, wb.z, acc[u][6]); acc[u][7] = fmaf(xv, wb.w, acc[u][7]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < C1_PX; ++u) {
            const size_t p = p0 + (size_t)u * T;
            if (p >= npix) continue;
            float* dst = y + p * (size_t)ldy;
            float o8[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) o8[o] = dact(acc[u][o], act);
            if (vec_out) {
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o8[0]), "f"(o8[1]), "f"(o8[2]),
                             "f"(o8[3]), "f"(o8[4]), "f"(o8[5]), "f"(o8[6]), "f"(o8[7])
                             : "memory");
            } else {
#pragma unroll
                for (int o = 0; o < 8; ++o)
                    if (o < Cout) dst[o] = o8[o];
            }
        }
    }
}

// z: the nine tap maps (row-major taps: t = dy * 3 + dx), either NHWC [B, Hi, Wi, ldz] (channels 0..8) or, with ldz == 0, PLANAR
// [B, 9, Hi, Wi] -- adjacent output pixels then read adjacent floats of one plane (the NHWC form touches one 32-byte sector per
// lane and tap for 4 useful bytes; whole occlusion tail 0.97 -> 0.84 ms at 1088x1920, batch 8); y [B, 2*Hi, 2*Wi].
// Source coordinates exactly as ATen's bilinear upsampling with align_corners=True (area_pixel_compute_source_index):
// the same expressions as resize_bilinear_nhwc_kernel.
// Tile = 32 x 16 outputs per CTA (two rows per thread).  The nine tap maps of the source rows / columns the tile can reach
// ((16 + 2) / 2 + 2 rows, (32 + 2) / 2 + 2 columns) are staged in shared memory once; the 36 samples per output then come from there
// (the first version read them from global memory: 36 L1 requests with 64-bit address arithmetic per pixel, 0.64 ms per call at
// 1088x1920, batch 8).  Same expressions and tap order as before: bit-identical.
constexpr int TS_TW = 32, TS_TH = 16, TS_RH = 12, TS_RW = 20;
__global__ void __launch_bounds__(256) upsample2_tapsum_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                                                               float* __restrict__ y, int Hi, int Wi, int ldz, float sy, float sx,
                                                               int act) {
    __shared__ float zs[9][TS_RH][TS_RW];
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    const int ox0 = blockIdx.x * TS_TW, oy0 = blockIdx.y * TS_TH;
    const size_t zplane = (size_t)Hi * Wi;
    const bool planar = (ldz == 0);
    const float* Z = z + (size_t)blockIdx.z * zplane * (planar ? 9 : ldz);
    const size_t tstride = planar ? zplane : 1, pstride = planar ? 1 : (size_t)ldz;
    // source rows / columns reached by the padded tile [oy0 - 1, oy0 + TS_TH] x [ox0 - 1, ox0 + TS_TW] (clipped to the image)
    const int ry0 = min((int)(sy * (float)max(oy0 - 1, 0)), Hi - 1), rx0 = min((int)(sx * (float)max(ox0 - 1, 0)), Wi - 1);
    const int ry1 = min(min((int)(sy * (float)min(oy0 + TS_TH, Ho - 1)), Hi - 1) + 1, Hi - 1);
    const int rx1 = min(min((int)(sx * (float)min(ox0 + TS_TW, Wo - 1)), Wi - 1) + 1, Wi - 1);
    const int rh = ry1 - ry0 + 1, rw = rx1 - rx0 + 1;              // <= TS_RH, TS_RW for scale factors <= 1/2 (checked on the host)
    for (int q = threadIdx.x; q < 9 * rh * rw; q += blockDim.x) {
        const int t = q / (rh * rw), r = q - t * (rh * rw), yy = r / rw, xx = r - yy * rw;
        zs[t][yy][xx] = __ldg(Z + t * tstride + ((size_t)(ry0 + yy) * Wi + rx0 + xx) * pstride);
    }
    __syncthreads();
    const int ox = ox0 + (threadIdx.x & 31);
    if (ox >= Wo) return;
    const float b0 = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int oy = oy0 + (threadIdx.x >> 5) + half * 8;
        if (oy >= Ho) continue;
        float acc = b0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int py = oy + dy - 1;
            if (py < 0 || py >= Ho) continue;                         // zero padding of the 3x3 convolution
            const float fy = sy * py;
            const int y0 = min((int)fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
            const float ly = fy - (float)y0, hy = 1.f - ly;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int px = ox + dx - 1;
                if (px < 0 || px >= Wo) continue;
                const float fx = sx * px;
                const int x0 = min((int)fx, Wi - 1), x1 = min(x0 + 1, Wi - 1);
                const float lx = fx - (float)x0, hx = 1.f - lx;
                const float(*Zt)[TS_RW] = zs[dy * 3 + dx];
                const float a = Zt[y0 - ry0][x0 - rx0], b = Zt[y0 - ry0][x1 - rx0];
                const float c = Zt[y1 - ry0][x0 - rx0], d = Zt[y1 - ry0][x1 - rx0];
                acc += hy * (hx * a + lx * b) + ly * (hx * c + lx * d);
            }
        }
        y[((size_t)blockIdx.z * Ho + oy) * Wo + ox] = dact(acc, act);
    }
}

}  // namespace fvfi

using namespace fvfi;

extern "C" int fvfi_conv1x1_nhwc(const float* x, int x_pixel_stride, const float* weight, const float* bias, float* y,
                                 int y_pixel_stride, size_t npix, int Cin, int Cout, int activation, void* stream) {
    FVFI_CHECK_ARG(x && weight && y && npix > 0, "conv1x1: bad argument");
    FVFI_CHECK_ARG(Cout >= 1 && Cout <= 8, "conv1x1: the direct kernel takes Cout <= 8 (got %d); use fvfi_conv2d_nhwc", Cout);
    FVFI_CHECK_ARG(Cin >= 8 && Cin <= 128 && (Cin & 7) == 0, "conv1x1: Cin must be a multiple of 8 in [8, 128] (got %d)", Cin);
    FVFI_CHECK_ARG((x_pixel_stride & 7) == 0 && x_pixel_stride >= Cin && ((((size_t)x) & 31) == 0),
                   "conv1x1: input pixels must be 32-byte aligned (pixel stride %d)", x_pixel_stride);
    FVFI_CHECK_ARG(y_pixel_stride >= Cout, "conv1x1: output pixel stride smaller than Cout");
    FVFI_CHECK_ARG(activation >= DACT_NONE && activation <= DACT_SIGMOID, "conv1x1: bad activation");
    const int nsm = sm_count() > 0 ? sm_count() : 148;
    const size_t want = (npix + (size_t)C1_THREADS * C1_PX - 1) / ((size_t)C1_THREADS * C1_PX);
    const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)nsm * 8 * 4);      // a few waves of 8 CTAs / SM
    cudaStream_t s = (cudaStream_t)stream;
#define FVFI_C1(N)                                                                                                             \
    case N:                                                                                                                    \
        conv1x1_small_kernel<N><<<grid, C1_THREADS, 0, s>>>(x, weight, bias, y, npix, Cout, x_pixel_stride, y_pixel_stride,      \
                                                            activation);                                                       \
        break;
    switch (Cin / 8) {
        FVFI_C1(1) FVFI_C1(2) FVFI_C1(3) FVFI_C1(4) FVFI_C1(5) FVFI_C1(6) FVFI_C1(7) FVFI_C1(8)
        FVFI_C1(9) FVFI_C1(10) FVFI_C1(11) FVFI_C1(12) FVFI_C1(13) FVFI_C1(14) FVFI_C1(15) FVFI_C1(16)
    }
#undef FVFI_C1
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

extern "C" int fvfi_upsample2_tapsum(const float* z, int z_pixel_stride, const float* bias, float* y, int B, int Hi, int Wi,
                                     int activation, void* stream) {
    FVFI_CHECK_ARG(z && y && B > 0 && B <= 65535 && Hi > 0 && Wi > 0, "upsample2_tapsum: bad argument");
    FVFI_CHECK_ARG(z_pixel_stride >= 9 || z_pixel_stride == 0,
                   "upsample2_tapsum: needs the nine tap channels per pixel (pixel stride %d; 0 = planar [B,9,Hi,Wi])", z_pixel_stride);
    FVFI_CHECK_ARG(activation >= DACT_NONE && activation <= DACT_SIGMOID, "upsample2_tapsum: bad activation");
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    dim3 grid((unsigned)ceil_div(Wo, TS_TW), (unsigned)ceil_div(Ho, TS_TH), (unsigned)B);
    FVFI_CHECK_ARG(grid.y <= 65535, "upsample2_tapsum: image too tall");
    upsample2_tapsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, bias, y, Hi, Wi, z_pixel_stride, sy, sx, activation);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}
