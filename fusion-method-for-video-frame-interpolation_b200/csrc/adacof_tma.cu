// adacof_tma.cu -- AdaCoF warp (forward, fused two-frame synthesis) with TMA-streamed coefficient maps.
//
// The coefficient maps (weight, alpha, beta: 3 x 25 planes, 300 B per output pixel) are the HBM-dominant stream of
// the warp (src/adacof/cupy_module/adacof.py:6-65).  In adacof_tiled.cu every thread prefetches its own coefficients one
// tap-row ahead in registers; the bytes in flight per SM are then bounded by registers x resident warps (~45 KB), which caps
// the kernel at 0.47-0.60 of the HBM roofline (ncu: long-scoreboard stalls, 55 % issue utilisation).  Here:
//   * persistent CTAs (2 per SM) walk 32 x 8 output tiles; a PRODUCER warp streams, per tile and tap-row, the 3 x 5 coefficient
//     planes of the tile with three cp.async.bulk.tensor (TMA, 3-D boxes 32 x 8 x 5, out-of-image elements zero-filled) into a
//     ring of shared-memory stages guarded by full/empty mbarriers -- bytes in flight no longer cost registers or warps;
//   * the 8 CONSUMER warps (one pixel per thread) read their coefficients from the stage (conflict-free LDS.32: a lane is a
//     column), gather the frame from the staged region (PLANAR here: the shared-memory pipe is the limiter once HBM latency
//     is hidden, and 12 LDS.32 cost 12 wavefronts where 4 LDS.128 of padded {R,G,B,-} pixels cost 16; clamped fill, global
//     fallback for offsets beyond the halo as in the tiled kernel), and release the stage;
//   * the frame region of the NEXT work item is fetched with 4-byte cp.async into the second region buffer while the current
//     item is processed, so tiles (and the two frames of the fused synthesis) follow each other without a staging bubble.
// Arithmetic (tap order, truncation toward zero, clamp-to-edge, fmaf contraction) is identical to adacof_tiled.cu, which stays
// the path for other filter sizes / dilations / unaligned widths.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace fvfi {

constexpr int XW = 32, XH = 8;                 // output tile
constexpr int XCONS = XW * XH;                 // consumer threads (one pixel each)
constexpr int XTHREADS = XCONS + 32;           // + producer warp
constexpr int XF = 5, XPADF = XF - 1, XHALO = 8;
constexpr int XSH = XH + XPADF + 2 * XHALO + 1;    // staged region rows / columns (tile + taps + halo + 1 for the +1 neighbour)
constexpr int XSW = XW + XPADF + 2 * XHALO + 1;
constexpr int XSTAGE_FLOATS = 3 * XF * XH * XW;    // one tap-row of the three maps
constexpr int XRPLANE = XSH * XSW;                 // one colour plane of the staged region
constexpr size_t XREGION_BYTES = ((size_t)3 * XRPLANE * sizeof(float) + 127) & ~(size_t)127;
constexpr unsigned XSPIN_LIMIT = 400u * 1000u * 1000u;

struct TmaArgs {
    const float* in[2];
    const float* occ;
    float* t[2];
    float* frame;
    float* mask;
    const float* gout;         // backward (NFRAMES == 0): gradOutput and the three gradient maps
    float* gw;
    float* goi;
    float* goj;
    int Hin, Win, H, W;
    int tiles_x, tiles_y, ntiles;
    int out_rows;              // rows of the OUTPUT planes t / frame / mask (<= H: the /32 padding rows are not written), plane = out_rows * W
};

__device__ __forceinline__ unsigned xs_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xs_u32(bar)), "r"(count));
}
__device__ __forceinline__ void xbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xs_u32(bar)) : "memory");
}
__device__ __forceinline__ void xbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xs_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = xs_u32(bar);
    unsigned done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > XSPIN_LIMIT) __trap();     // bounded wait: a protocol bug traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(xs_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(xs_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(xs_u32(dst)), "l"(src) : "memory");
}

struct XTile { int n, i0, j0; };
__device__ __forceinline__ XTile x_tile(const TmaArgs& A, int tile) {
    XTile t;
    const int per = A.tiles_x * A.tiles_y;
    t.n = tile / per;
    const int r = tile - t.n * per;
    const int ty = r / A.tiles_x;
    t.i0 = ty * XH;
    t.j0 = (r - ty * A.tiles_x) * XW;
    return t;
}

// NFRAMES: 1 = forward, 2 = fused two-frame synthesis, 0 = fused backward (gW, g_alpha, g_beta from one gather; algebra in adacof.cu)
// STATS (NFRAMES == 2): accumulate the offset moments of both frames for the flow-variance mask; false when the caller passes no mask
// (the recipe's three baseline passes use the synthesised frame only, src/fusion_net/interpolate_twoframe.py:228-238)
template <int NFRAMES, int XSTAGES, int MINB, bool STATS = true>
__global__ void __launch_bounds__(XTHREADS, MINB)
adacof_fwd_tma(const __grid_constant__ CUtensorMap mw0, const __grid_constant__ CUtensorMap ma0,
               const __grid_constant__ CUtensorMap mb0, const __grid_constant__ CUtensorMap mw1,
               const __grid_constant__ CUtensorMap ma1, const __grid_constant__ CUtensorMap mb1, const TmaArgs A) {
    constexpr bool BWD = NFRAMES == 0;
    constexpr int NF = BWD ? 1 : NFRAMES;
    extern __shared__ __align__(128) unsigned char xsm[];
    constexpr size_t XRING_BYTES = (size_t)XSTAGES * XSTAGE_FLOATS * sizeof(float);
    float* ring = (float*)xsm;
    float* region0 = (float*)(xsm + XRING_BYTES);
    float* region1 = (float*)(xsm + XRING_BYTES + XREGION_BYTES);
    unsigned long long* full = (unsigned long long*)(xsm + XRING_BYTES + 2 * XREGION_BYTES);
    unsigned long long* empty = full + XSTAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = (size_t)A.H * A.W, plane_in = (size_t)A.Hin * A.Win, oplane = (size_t)A.out_rows * A.W;
    if (threadIdx.x == 0) {
        for (int s = 0; s < XSTAGES; ++s) { xbar_init(&full[s], 1); xbar_init(&empty[s], XH); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == XH) {
        // ================= producer: one lane streams the coefficient tap-rows of every work item =================
        if (lane == 0) {
            unsigned cnt = 0;
            for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
                const XTile T = x_tile(A, tile);
#pragma unroll 1
                for (int f = 0; f < NF; ++f) {
                    const CUtensorMap* pw = f ? &mw1 : &mw0;
                    const CUtensorMap* pa = f ? &ma1 : &ma0;
                    const CUtensorMap* pb = f ? &mb1 : &mb0;
#pragma unroll 1
                    for (int k = 0; k < XF; ++k, ++cnt) {
                        const unsigned s = cnt % XSTAGES;
                        if (cnt >= XSTAGES) xbar_wait(&empty[s], ((cnt / XSTAGES) - 1) & 1);
                        xbar_expect_tx(&full[s], XSTAGE_FLOATS * sizeof(float));
                        float* st = ring + (size_t)s * XSTAGE_FLOATS;
                        const int z = T.n * XF * XF + k * XF;
                        tma_load_3d(st, pw, T.j0, T.i0, z, &full[s]);
                        tma_load_3d(st + XF * XH * XW, pa, T.j0, T.i0, z, &full[s]);
                        tma_load_3d(st + 2 * XF * XH * XW, pb, T.j0, T.i0, z, &full[s]);
                    }
                }
            }
        }
        return;
    }

    // ================= consumers: thread = pixel (row = warp, column = lane) of the tile =================
    auto issue_region = [&](float* R, const XTile& T, int f) {
        const float* I = A.in[f] + (size_t)T.n * 3 * plane_in;
        for (int p = threadIdx.x; p < XRPLANE; p += XCONS) {
            const int r = p / XSW, c = p - r * XSW;
            const int gr = min(max(T.i0 - XHALO + r, 0), A.Hin - 1);
            const int gc = min(max(T.j0 - XHALO + c, 0), A.Win - 1);
            const float* src = I + (size_t)gr * A.Win + gc;
            cp_async4(R + p, src);                                   // planar: consecutive lanes, consecutive words
            cp_async4(R + XRPLANE + p, src + plane_in);
            cp_async4(R + 2 * XRPLANE + p, src + 2 * plane_in);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    unsigned cnt = 0;
    int item = 0;
    if ((int)blockIdx.x < A.ntiles) issue_region(region0, x_tile(A, blockIdx.x), 0);
    float keep0 = 0.f, keep1 = 0.f, keep2 = 0.f, keepv = 0.f;      // frame-0 result of the fused synthesis
    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
        const XTile T = x_tile(A, tile);
        const int gi = T.i0 + warp, gj = T.j0 + lane;
        const bool live = gi < A.H && gj < A.W;
        const size_t p = (size_t)gi * A.W + gj;
#pragma unroll 1
        for (int f = 0; f < NF; ++f, ++item) {
            // this item's region has landed (own copies), everybody's copies have landed and everybody is done with the
            // previous item -> its region buffer may be refilled with the NEXT item's region
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(XCONS) : "memory");
            {
                const bool more_f = f + 1 < NF;
                const int ntile = more_f ? tile : tile + (int)gridDim.x;
                if (ntile < A.ntiles) issue_region((item & 1) ? region0 : region1, x_tile(A, ntile), more_f ? f + 1 : 0);
            }
            const float* R = (item & 1) ? region1 : region0;
            const float* I = A.in[f] + (size_t)T.n * 3 * plane_in;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (BWD && live) {
                g0 = ld_stream(A.gout + ((size_t)T.n * 3 + 0) * plane + p);
                g1 = ld_stream(A.gout + ((size_t)T.n * 3 + 1) * plane + p);
                g2 = ld_stream(A.gout + ((size_t)T.n * 3 + 2) * plane + p);
            }
            const size_t gq = (size_t)T.n * XF * XF * plane + p;      // this pixel in the [B,25,H,W] gradient maps
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
            float s0 = 0.f, s1i = 0.f, s2i = 0.f, s1j = 0.f, s2j = 0.f;
#pragma unroll 1
            for (int k = 0; k < XF; ++k, ++cnt) {
                const unsigned s = cnt % XSTAGES;
                xbar_wait(&full[s], (cnt / XSTAGES) & 1);
                const float* st = ring + (size_t)s * XSTAGE_FLOATS + warp * XW + lane;
#pragma unroll
                for (int l = 0; l < XF; ++l) {
                    const float w = st[l * XH * XW];
                    const float al = st[(XF + l) * XH * XW];
                    const float be = st[(2 * XF + l) * XH * XW];
                    const int Ai = (int)al, Bj = (int)be;                 // trunc toward zero -- adacof.py:27-28
                    const float a = al - (float)Ai, b = be - (float)Bj;
                    const float na = 1.f - a, nb = 1.f - b;
                    const int rr = warp + k + Ai + XHALO, cc = lane + l + Bj + XHALO;
                    float4 v00, v01, v10, v11;
                    if ((unsigned)rr < (unsigned)(XSH - 1) && (unsigned)cc < (unsigned)(XSW - 1)) {
                        // planar region: 12 LDS.32 (one wavefront each for smooth offsets) instead of 4 LDS.128 (four each)
                        const float* q = R + rr * XSW + cc;
                        v00 = make_float4(q[0], q[XRPLANE], q[2 * XRPLANE], 0.f);
                        v01 = make_float4(q[1], q[XRPLANE + 1], q[2 * XRPLANE + 1], 0.f);
                        v10 = make_float4(q[XSW], q[XRPLANE + XSW], q[2 * XRPLANE + XSW], 0.f);
                        v11 = make_float4(q[XSW + 1], q[XRPLANE + XSW + 1], q[2 * XRPLANE + XSW + 1], 0.f);
                    } else {                                              // offset beyond the halo: global, explicit clamps
                        const int gr = gi + k + Ai, gc = gj + l + Bj;
                        const int r0 = min(max(gr, 0), A.Hin - 1), r1 = min(max(gr + 1, 0), A.Hin - 1);
                        const int c0 = min(max(gc, 0), A.Win - 1), c1 = min(max(gc + 1, 0), A.Win - 1);
                        const float* p00 = I + (size_t)r0 * A.Win + c0;
                        const float* p10 = I + (size_t)r1 * A.Win + c0;
                        const float* p01 = I + (size_t)r0 * A.Win + c1;
                        const float* p11 = I + (size_t)r1 * A.Win + c1;
                        v00 = make_float4(__ldg(p00), __ldg(p00 + plane_in), __ldg(p00 + 2 * plane_in), 0.f);
                        v10 = make_float4(__ldg(p10), __ldg(p10 + plane_in), __ldg(p10 + 2 * plane_in), 0.f);
                        v01 = make_float4(__ldg(p01), __ldg(p01 + plane_in), __ldg(p01 + 2 * plane_in), 0.f);
                        v11 = make_float4(__ldg(p11), __ldg(p11 + plane_in), __ldg(p11 + 2 * plane_in), 0.f);
                    }
                    if (BWD) {
                        const float s00 = fmaf(g2, v00.z, fmaf(g1, v00.y, g0 * v00.x));
                        const float s10 = fmaf(g2, v10.z, fmaf(g1, v10.y, g0 * v10.x));
                        const float s01 = fmaf(g2, v01.z, fmaf(g1, v01.y, g0 * v01.x));
                        const float s11 = fmaf(g2, v11.z, fmaf(g1, v11.y, g0 * v11.x));
                        if (live) {
                            const size_t o = gq + (size_t)(k * XF + l) * plane;
                            st_stream(A.gw + o, s00 * (na * nb) + s10 * (a * nb) + s01 * (na * b) + s11 * (a * b));
                            st_stream(A.goi + o, w * ((s10 - s00) * nb + (s11 - s01) * b));
                            st_stream(A.goj + o, w * ((s01 - s00) * na + (s11 - s10) * a));
                        }
                    } else {
                        const float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
                        acc0 = fmaf(w, v00.x * w00 + v10.x * w10 + v01.x * w01 + v11.x * w11, acc0);
                        acc1 = fmaf(w, v00.y * w00 + v10.y * w10 + v01.y * w01 + v11.y * w11, acc1);
                        acc2 = fmaf(w, v00.z * w00 + v10.z * w10 + v01.z * w01 + v11.z * w11, acc2);
                    }
                    if (NFRAMES == 2 && STATS) {
                        s0 += w;
                        s1i = fmaf(w, al, s1i);
                        s2i = fmaf(w * al, al, s2i);
                        s1j = fmaf(w, be, s1j);
                        s2j = fmaf(w * be, be, s2j);
                    }
                }
                __syncwarp();
                if (lane == 0) xbar_arrive(&empty[s]);          // this warp is done with the stage
            }
            const bool live_out = live && gi < A.out_rows;        // p = gi * W + gj also indexes the (shorter) output planes
            if (!BWD && live_out) {
                float* const t = A.t[f];
                if (t) {
                    st_stream(t + ((size_t)T.n * 3 + 0) * oplane + p, acc0);
                    st_stream(t + ((size_t)T.n * 3 + 1) * oplane + p, acc1);
                    st_stream(t + ((size_t)T.n * 3 + 2) * oplane + p, acc2);
                }
            }
            if (NFRAMES == 2) {
                const float var = (s2i - s1i * s1i * (2.f - s0)) + (s2j - s1j * s1j * (2.f - s0));
                if (f == 0) {
                    keep0 = acc0; keep1 = acc1; keep2 = acc2; keepv = var;
                } else if (live_out) {
                    if (A.frame) {  // fusion_adacofnet.py:198
                        const float o = ld_stream(A.occ + (size_t)T.n * plane + p);
                        st_stream(A.frame + ((size_t)T.n * 3 + 0) * oplane + p, o * keep0 + (1.f - o) * acc0);
                        st_stream(A.frame + ((size_t)T.n * 3 + 1) * oplane + p, o * keep1 + (1.f - o) * acc1);
                        st_stream(A.frame + ((size_t)T.n * 3 + 2) * oplane + p, o * keep2 + (1.f - o) * acc2);
                    }
                    if (A.mask)  // fusion_adacofnet.py:211-213
                        st_stream(A.mask + (size_t)T.n * oplane + p, fminf(fmaxf(fmaxf(keepv, var), 0.f), 20.f) / 20.f);
                }
            }
        }
    }
}

// ---- host: tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// [B*25, H, W] fp32 planes -> 3-D map with 32 x 8 x 5 boxes
static bool make_map(CUtensorMap* m, const float* base, int B, int H, int W) {
    // the caching allocator hands the same buffers back call after call: remember the last few encodings per host thread
    struct Slot { const float* base; int B, H, W; CUtensorMap map; };
    static thread_local Slot cache[16];
    static thread_local unsigned next = 0;
    for (const Slot& c : cache)
        if (c.base == base && c.B == B && c.H == H && c.W == W) { *m = c.map; return true; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * XF * XF};
    const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)H * W * sizeof(float)};
    const cuuint32_t box[3] = {XW, XH, XF};
    const cuuint32_t es[3] = {1, 1, 1};
    if (enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    Slot& c = cache[next++ % 16];
    c.base = base; c.B = B; c.H = H; c.W = W; c.map = *m;
    return true;
}

static bool aligned16(const void* p) { return (((size_t)p) & 15) == 0; }

// Returns FVFI_OK with *handled = 1 if the TMA path ran, *handled = 0 if it does not apply (caller falls back).
// nframes: 1 forward, 2 fused synthesis, 0 backward (gout, gw/goi/goj via the bwd_* arguments)
int adacof_tma_launch(const float* in1, const float* in2, const float* w1, const float* a1, const float* b1,
                      const float* w2, const float* a2, const float* b2, const float* occ, float* t1, float* t2,
                      float* frame, float* mask, int nframes, int B, int Hin, int Win, int H, int W, int F, int dil,
                      cudaStream_t s, int* handled, const float* bwd_gout = nullptr, float* bwd_gw = nullptr,
                      float* bwd_goi = nullptr, float* bwd_goj = nullptr, int out_rows = 0) {
    *handled = 0;
    if (F != XF || dil != 1 || (W & 3) || !aligned16(w1) || !aligned16(a1) || !aligned16(b1)) return FVFI_OK;
    if (nframes == 2 && (!aligned16(w2) || !aligned16(a2) || !aligned16(b2))) return FVFI_OK;
    CUtensorMap mw0, ma0, mb0, mw1, ma1, mb1;
    if (!make_map(&mw0, w1, B, H, W) || !make_map(&ma0, a1, B, H, W) || !make_map(&mb0, b1, B, H, W)) return FVFI_OK;
    if (nframes == 2) {
        if (!make_map(&mw1, w2, B, H, W) || !make_map(&ma1, a2, B, H, W) || !make_map(&mb1, b2, B, H, W)) return FVFI_OK;
    } else {
        mw1 = mw0; ma1 = ma0; mb1 = mb0;
    }
    TmaArgs A{};
    A.in[0] = in1; A.in[1] = in2; A.occ = occ; A.t[0] = t1; A.t[1] = t2; A.frame = frame; A.mask = mask;
    A.gout = bwd_gout; A.gw = bwd_gw; A.goi = bwd_goi; A.goj = bwd_goj;
    A.Hin = Hin; A.Win = Win; A.H = H; A.W = W;
    A.out_rows = (out_rows > 0 && out_rows < H) ? out_rows : H;
    A.tiles_x = ceil_div(W, XW);
    A.tiles_y = ceil_div(H, XH);
    const long long nt = (long long)A.tiles_x * A.tiles_y * B;
    if (nt > 0x7fffffffLL) return FVFI_OK;
    A.ntiles = (int)nt;
    const int nsm = sm_count() > 0 ? sm_count() : 148;
#define FVFI_TMA_LAUNCH(NF, ST, MB, ...)                                                                                  \
    {                                                                                                                     \
        const size_t smem = (size_t)ST * XSTAGE_FLOATS * sizeof(float) + 2 * XREGION_BYTES + 128;                        \
        const unsigned grid = (unsigned)std::min<long long>(nt, (long long)MB * nsm);                                    \
        FVFI_SMEM_OPT_IN((adacof_fwd_tma<NF, ST, MB, ##__VA_ARGS__>), smem);                                              \
        adacof_fwd_tma<NF, ST, MB, ##__VA_ARGS__><<<grid, XTHREADS, smem, s>>>(mw0, ma0, mb0, mw1, ma1, mb1, A);        \
    }
    if (nframes == 0) FVFI_TMA_LAUNCH(0, 3, 2) else
    // measured (tools/prof_adacof.py): the single warp is fastest with a 3-deep ring and 2 CTAs/SM, the fused synthesis (more
    // arithmetic per byte: moments, blend) with a 2-deep ring and 3 CTAs/SM (67 KB each, 72 registers)
    if (nframes == 2 && !mask) FVFI_TMA_LAUNCH(2, 2, 3, false) else
    if (nframes == 2) FVFI_TMA_LAUNCH(2, 2, 3) else FVFI_TMA_LAUNCH(1, 3, 2)
#undef FVFI_TMA_LAUNCH
    FVFI_LAUNCH_CHECK();
    *handled = 1;
    return FVFI_OK;
}

}  // namespace fvfi
