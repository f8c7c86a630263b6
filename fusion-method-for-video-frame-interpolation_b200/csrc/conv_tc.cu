// conv_tc.cu -- implicit-GEMM 2-D convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fp32 in / fp32 out,
// 3xTF32 error-compensated so the result matches an fp32 FFMA convolution to ~1e-6 relative.
//
// Used for the PhaseNet / KernelEstimation / FusionNet convolutions (the only dense contractions on the path;
// reference: torch.nn.Conv2d -> cuDNN, src/phase_net/phase_net.py:190-199, src/fusion_net/fusion_adacofnet.py:19-83,
// src/fusion_net/fusion_net.py:24-36).  Plain TF32 (10-bit mantissa) does not hold the 1e-4 output bound through
// ~30 layers, so every product a*b is formed as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with a_hi = rna_tf32(a),
// a_lo = a - a_hi (exact), three tcgen05.mma.kind::tf32 per K-step accumulating in fp32 in TMEM.
//
// Formulation (stride 1, "same" padding, zero or reflect):
//   activations NHWC (torch channels_last), GEMM M = output pixels, N = Cout, K = taps x Cin.
//   A CTA owns an output patch of 16 rows x (8*MT) columns = MT accumulator tiles of M = 128 (16 rows x 8 px),
//   each tile N columns of TMEM.  Per 16-channel chunk the loader warps stage the (16+KH-1) x (8*MT+KW-1) input
//   region ONCE, split into hi/lo, in the no-swizzle K-major canonical layout [kchunk(16B)][pixel][16B]; every
//   filter tap is then just a different descriptor start address into that region (SBO = region row pitch), so
//   the activation is read from L2 once per chunk instead of once per tap.  Weights are pre-packed on the device
//   (hi|lo, canonical layout) and streamed per (chunk, tap) with cp.async.bulk + mbarrier.
//   Warp roles: warps 0-3 activation loaders + epilogue (TMEM -> regs -> bias/activation -> NHWC), warp 4 weight
//   producer, warp 5 TMEM allocator + single-thread MMA issuer.
#include "common.cuh"

namespace fvfi {

constexpr int CV_CHUNK = 16;            // input channels per K chunk (2 MMAs of K=8)
constexpr int CV_ROWS = 16;             // output rows per CTA
constexpr int CV_LOADERS = 128;         // warps 0-3
constexpr int CV_THREADS = 192;
constexpr int CV_ASTAGES = 2;
constexpr int CV_MAX_BSTAGES = 4;
constexpr unsigned CV_SPIN_LIMIT = 200u * 1000u * 1000u;   // bounded waits: trap instead of hanging the GPU

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };
enum { PAD_ZERO = 0, PAD_REFLECT = 1 };

struct ConvArgs {
    const float* x;        // [B,H,W,Cin] NHWC
    const float* wpack;    // packed weights (see pack kernel)
    const float* bias;     // [Npad] or null
    float* y;              // [B,H,W,Cout] NHWC
    int B, H, W, Cin, Cout, Npad, KH, KW, pad_mode, act;
    int ldx, ldy;          // floats per pixel in the input / output storage (channel-slice views)
    int MT, RW, RH, NPIX;  // tiles per CTA, staged region geometry
    int nchunks, bstages, tmem_cols;
    unsigned a_stage_bytes, b_stage_bytes;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
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
        if (++spins > CV_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(unsigned taddr, float* v) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (between the two 16 B K-chunks of one MMA) |
//   [32,46) SBO >> 4 (between 8-row core matrices along M/N) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fffu);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ float to_tf32_rna(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_ELU: return v > 0.f ? v : expm1f(v);
        case ACT_TANH: return tanhf(v);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {   // torch 'reflect': -1 -> 1, n -> n-2
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

// ---- weight packing: OIHW fp32 -> [chunk][tap][hi|lo][kc(4)][n(Npad)][4] --------------------------------------
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KH,
                                         int KW, int Npad, int nchunks) {
    const int taps = KH * KW;
    const size_t total = (size_t)nchunks * taps * 2 * 4 * Npad * 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        size_t t = q;
        const int e = (int)(t % 4); t /= 4;
        const int n = (int)(t % Npad); t /= Npad;
        const int kc = (int)(t % 4); t /= 4;
        const int lo = (int)(t % 2); t /= 2;
        const int tap = (int)(t % taps); t /= taps;
        const int chunk = (int)t;
        const int c = chunk * CV_CHUNK + kc * 4 + e;
        float v = 0.f;
        if (n < Cout && c < Cin) v = w[(((size_t)n * Cin + c) * KH + tap / KW) * KW + tap % KW];
        unsigned hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        out[q] = lo ? (v - hi) : hi;
    }
}

// ---- the convolution ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_THREADS, 1) conv_tf32x3_kernel(const ConvArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [A stages: hi, lo] [B stages] [barriers] [tmem ptr]
    unsigned char* a_base = smem_raw;
    unsigned char* b_base = a_base + (size_t)CV_ASTAGES * A.a_stage_bytes;
    unsigned long long* bars = (unsigned long long*)(b_base + (size_t)A.bstages * A.b_stage_bytes);
    unsigned long long* a_full = bars;                       // [2]  count 128
    unsigned long long* a_empty = bars + 2;                  // [2]  count 1 (tcgen05.commit)
    unsigned long long* b_full = bars + 4;                   // [bstages] count 1 + tx
    unsigned long long* b_empty = bars + 4 + CV_MAX_BSTAGES; // [bstages] count 1 (tcgen05.commit)
    unsigned long long* acc_full = bars + 4 + 2 * CV_MAX_BSTAGES;
    unsigned* tmem_ptr = (unsigned*)(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = A.KH * A.KW;
    const int x0 = blockIdx.x * 8 * A.MT, y0 = blockIdx.y * CV_ROWS, img = blockIdx.z;
    const unsigned a_half = A.a_stage_bytes / 2;             // hi | lo halves of an A stage

    if (threadIdx.x == 0) {
        for (int s = 0; s < CV_ASTAGES; ++s) { mbar_init(&a_full[s], CV_LOADERS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < A.bstages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {   // TMEM allocation by one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(A.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_ptr;

    if (warp < 4) {
        // ================= activation loaders: region -> (hi, lo) canonical tiles =================
        const int padT = A.KH / 2, padL = A.KW / 2;
        const float* X = A.x + (size_t)img * A.H * A.W * A.ldx;
        const bool vec = ((A.ldx & 3) == 0) && ((((size_t)A.x) & 15) == 0);
        for (int c = 0; c < A.nchunks; ++c) {
            const int s = c % CV_ASTAGES;
            if (c >= CV_ASTAGES) mbar_wait(&a_empty[s], ((c / CV_ASTAGES) - 1) & 1);
            float4* hi = (float4*)(a_base + (size_t)s * A.a_stage_bytes);
            float4* lo = (float4*)(a_base + (size_t)s * A.a_stage_bytes + a_half);
            const int cbase = c * CV_CHUNK;
            for (int q = threadIdx.x; q < A.NPIX * 4; q += CV_LOADERS) {
                const int pix = q >> 2, kc = q & 3;              // 4 consecutive threads read one pixel's 64 B
                const int r = pix / A.RW, cc = pix - r * A.RW;
                int gy = y0 + r - padT, gx = x0 + cc - padL;
                bool ok = true;
                if (A.pad_mode == PAD_REFLECT) {
                    gy = reflect101(gy, A.H);
                    gx = reflect101(gx, A.W);
                } else {
                    ok = (gy >= 0 && gy < A.H && gx >= 0 && gx < A.W);
                }
                const int ch = cbase + kc * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && ch < A.Cin) {
                    const float* p = X + ((size_t)gy * A.W + gx) * A.ldx + ch;
                    if (vec && ch + 3 < A.Cin) {
                        v = __ldg((const float4*)p);
                    } else {
                        v.x = __ldg(p);
                        if (ch + 1 < A.Cin) v.y = __ldg(p + 1);
                        if (ch + 2 < A.Cin) v.z = __ldg(p + 2);
                        if (ch + 3 < A.Cin) v.w = __ldg(p + 3);
                    }
                }
                float4 h;
                h.x = to_tf32_rna(v.x); h.y = to_tf32_rna(v.y); h.z = to_tf32_rna(v.z); h.w = to_tf32_rna(v.w);
                const int o = kc * A.NPIX + pix;
                hi[o] = h;
                lo[o] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            fence_async_smem();          // generic-proxy stores -> visible to the tensor-core (async) proxy
            mbar_arrive(&a_full[s]);
        }
        // ================= epilogue =================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int m = warp * 32 + lane;                          // accumulator row = TMEM lane
        const int orow = y0 + (m >> 3);
        for (int t = 0; t < A.MT; ++t) {
            const int ocol = x0 + t * 8 + (m & 7);
            const bool inb = (orow < A.H && ocol < A.W);
            float* dst = A.y + (((size_t)img * A.H + orow) * A.W + ocol) * A.ldy;
            for (int n0 = 0; n0 < A.Npad; n0 += 16) {
                float v[16];
                tc_ld16(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(t * A.Npad + n0), v);
                if (!inb) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = n0 + i;
                    if (n < A.Cout) v[i] = apply_act(v[i] + (A.bias ? __ldg(A.bias + n) : 0.f), A.act);
                }
                if ((A.Cout & 3) == 0 && (A.ldy & 3) == 0 && ((((size_t)A.y) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (n0 + i < A.Cout) *(float4*)(dst + n0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < A.Cout) dst[n0 + i] = v[i];
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ================= weight producer =================
        if (lane == 0) {
            int it = 0;
            for (int c = 0; c < A.nchunks; ++c)
                for (int tp = 0; tp < taps; ++tp, ++it) {
                    const int s = it % A.bstages;
                    if (it >= A.bstages) mbar_wait(&b_empty[s], ((it / A.bstages) - 1) & 1);
                    mbar_expect_tx(&b_full[s], A.b_stage_bytes);
                    bulk_g2s(b_base + (size_t)s * A.b_stage_bytes,
                             A.wpack + ((size_t)c * taps + tp) * (A.b_stage_bytes / 4), A.b_stage_bytes, &b_full[s]);
                }
        }
    } else {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=TF32, K-major both, N = Npad, M = 128
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(A.Npad >> 3) << 17) | ((128u >> 4) << 24);
            const unsigned a_lbo = (unsigned)A.NPIX * 16u, a_sbo = (unsigned)A.RW * 16u;
            const unsigned b_lbo = (unsigned)A.Npad * 16u, b_sbo = 128u;
            const unsigned b_half = A.b_stage_bytes / 2;
            int it = 0;
            for (int c = 0; c < A.nchunks; ++c) {
                const int sa = c % CV_ASTAGES;
                mbar_wait(&a_full[sa], (c / CV_ASTAGES) & 1);
                const unsigned a_hi = smem_u32(a_base + (size_t)sa * A.a_stage_bytes), a_lo = a_hi + a_half;
                for (int tp = 0; tp < taps; ++tp, ++it) {
                    const int sb = it % A.bstages;
                    mbar_wait(&b_full[sb], (it / A.bstages) & 1);
                    tc_fence_after();
                    const unsigned b_hi = smem_u32(b_base + (size_t)sb * A.b_stage_bytes), b_lo = b_hi + b_half;
                    const int dy = tp / A.KW, dx = tp - dy * A.KW;
                    for (int t = 0; t < A.MT; ++t) {
                        const unsigned pix_off = (unsigned)(dy * A.RW + dx + 8 * t) * 16u;
                        const unsigned d = tmem + (unsigned)(t * A.Npad);
                        for (int ks = 0; ks < 2; ++ks) {   // two K=8 steps per 16-channel chunk
                            const unsigned ak = (unsigned)(2 * ks) * a_lbo + pix_off;
                            const unsigned bk = (unsigned)(2 * ks) * b_lbo;
                            const unsigned long long dah = make_desc(a_hi + ak, a_lbo, a_sbo);
                            const unsigned long long dal = make_desc(a_lo + ak, a_lbo, a_sbo);
                            const unsigned long long dbh = make_desc(b_hi + bk, b_lbo, b_sbo);
                            const unsigned long long dbl = make_desc(b_lo + bk, b_lbo, b_sbo);
                            const unsigned acc0 = (c == 0 && tp == 0 && ks == 0) ? 0u : 1u;   // first MMA of this tile
                            tc_mma_tf32(d, dal, dbh, idesc, acc0);      // small terms first
                            tc_mma_tf32(d, dah, dbl, idesc, 1u);
                            tc_mma_tf32(d, dah, dbh, idesc, 1u);
                        }
                    }
                    tc_commit(&b_empty[sb]);           // weights of this (chunk, tap) consumed
                }
                tc_commit(&a_empty[sa]);               // region of this chunk consumed
            }
            tc_commit(acc_full);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(A.tmem_cols));
    }
}

// ---- host side -----------------------------------------------------------------------------------
static int conv_geometry(ConvArgs& a, size_t* smem_bytes) {
    a.Npad = (a.Cout + 15) & ~15;
    FVFI_CHECK_ARG(a.Npad >= 16 && a.Npad <= 256, "conv: Cout %d not supported (1..256 per launch)", a.Cout);
    a.nchunks = (a.Cin + CV_CHUNK - 1) / CV_CHUNK;
    a.b_stage_bytes = (unsigned)a.Npad * CV_CHUNK * 4u * 2u;           // hi + lo
    const size_t budget = 220 * 1024;
    for (int mt = 4; mt >= 1; mt >>= 1) {
        if (mt * a.Npad > 512) continue;
        if (mt > 1 && 8 * (mt / 2) >= a.W) continue;                   // do not over-tile narrow images
        a.MT = mt;
        a.RW = 8 * mt + a.KW - 1;
        a.RH = CV_ROWS + a.KH - 1;
        a.NPIX = a.RW * a.RH;
        a.a_stage_bytes = (unsigned)a.NPIX * CV_CHUNK * 4u * 2u;       // hi + lo
        const size_t fixed = (size_t)CV_ASTAGES * a.a_stage_bytes + 256;
        if (fixed + 2 * (size_t)a.b_stage_bytes > budget) continue;
        int bs = (int)((budget - fixed) / a.b_stage_bytes);
        a.bstages = bs > CV_MAX_BSTAGES ? CV_MAX_BSTAGES : bs;
        int cols = 32;
        while (cols < mt * a.Npad) cols <<= 1;
        a.tmem_cols = cols;
        *smem_bytes = fixed + (size_t)a.bstages * a.b_stage_bytes;
        return FVFI_OK;
    }
    set_error("conv: no tile configuration fits (Cout %d, kernel %dx%d)", a.Cout, a.KH, a.KW);
    return FVFI_EINVAL;
}

}  // namespace fvfi

using namespace fvfi;

extern "C" size_t fvfi_conv2d_packed_weight_floats(int Cout, int Cin, int KH, int KW) {
    const int Npad = (Cout + 15) & ~15;
    const int nchunks = (Cin + CV_CHUNK - 1) / CV_CHUNK;
    return (size_t)nchunks * KH * KW * 2 * 4 * Npad * 4;
}

extern "C" int fvfi_conv2d_pack_weights(const float* weight_oihw, float* packed, int Cout, int Cin, int KH, int KW,
                                        void* stream) {
    FVFI_CHECK_ARG(weight_oihw && packed && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "conv_pack_weights: bad argument");
    const int Npad = (Cout + 15) & ~15;
    const int nchunks = (Cin + CV_CHUNK - 1) / CV_CHUNK;
    const size_t total = fvfi_conv2d_packed_weight_floats(Cout, Cin, KH, KW);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    conv_pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(weight_oihw, packed, Cout, Cin, KH, KW, Npad, nchunks);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}

extern "C" int fvfi_conv2d_nhwc(const float* x, int x_pixel_stride, const float* packed_weight, const float* bias, float* y,
                                int y_pixel_stride, int B, int H, int W, int Cin, int Cout, int KH, int KW, int pad_mode,
                                int activation, void* stream) {
    FVFI_CHECK_ARG(x && packed_weight && y, "conv2d: null pointer");
    FVFI_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && B <= 65535, "conv2d: bad dimension");
    FVFI_CHECK_ARG((KH == 1 || KH == 3 || KH == 5) && KW == KH, "conv2d: kernel must be 1x1, 3x3 or 5x5");
    FVFI_CHECK_ARG(pad_mode == PAD_ZERO || pad_mode == PAD_REFLECT, "conv2d: pad_mode must be 0 (zero) or 1 (reflect)");
    FVFI_CHECK_ARG(pad_mode != PAD_REFLECT || (H > KH / 2 && W > KW / 2), "conv2d: reflect padding needs H,W > pad");
    FVFI_CHECK_ARG(activation >= 0 && activation <= 4, "conv2d: bad activation");
    ConvArgs a{};
    a.x = x; a.wpack = packed_weight; a.bias = bias; a.y = y; a.ldx = x_pixel_stride; a.ldy = y_pixel_stride;
    FVFI_CHECK_ARG(x_pixel_stride >= Cin && y_pixel_stride >= Cout, "conv2d: pixel stride smaller than channel count");
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.KH = KH; a.KW = KW; a.pad_mode = pad_mode; a.act = activation;
    size_t smem = 0;
    if (int rc = conv_geometry(a, &smem)) return rc;
    FVFI_CUDA(cudaFuncSetAttribute(conv_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(W, 8 * a.MT), ceil_div(H, CV_ROWS), B);
    conv_tf32x3_kernel<<<grid, CV_THREADS, smem, (cudaStream_t)stream>>>(a);
    FVFI_LAUNCH_CHECK();
    return FVFI_OK;
}
