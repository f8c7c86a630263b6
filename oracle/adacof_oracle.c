/*
 * oracle/adacof_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32) of the reference's AdaCoF warp arithmetic, used
 * only as the parity checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under
 * fusion-method-for-video-frame-interpolation_b200/ may link or call this file.
 *
 * The reference has NO CPU implementation of this operator
 * (src/adacof/cupy_module/adacof.py:356-357 raises NotImplementedError); this
 * file restates the four CUDA C kernels that live as Python strings in
 *   src/adacof/cupy_module/adacof.py:6-65    kernel_AdaCoF_updateOutput
 *   src/adacof/cupy_module/adacof.py:67-128  kernel_AdaCoF_updateGradWeight
 *   src/adacof/cupy_module/adacof.py:130-193 kernel_AdaCoF_updateGradAlpha
 *   src/adacof/cupy_module/adacof.py:195-258 kernel_AdaCoF_updateGradBeta
 * Pinning: the restatement is checked on a B200 against those very kernels,
 * expanded by the reference's own cupy_kernel() and compiled to cubins under
 * oracle/_ref/ by oracle/build_ref_kernels.py (tests/test_adacof_gpu.py).
 *
 * Semantics kept on purpose:
 *   - (int) cast = truncation toward zero, so negative fractional offsets give
 *     bilinear weights outside [0,1]            (adacof.py:27-28)
 *   - every tap coordinate is clamped to the padded frame (adacof.py:30-52)
 *   - tap order is k-major, fp32 accumulation   (adacof.py:22-23,54-59)
 *   - gradients sum over exactly 3 channels     (adacof.py:86,150,215)
 *   - gradInput is never computed (zeros)       (adacof.py:382,445)
 * Built with -ffp-contract=off so the expression tree is evaluated as written;
 * the CUDA kernels (NVRTC default --fmad=true) may differ by ~1 ulp per term,
 * which the parity tests cover with a stated tolerance.
 */
#include <stddef.h>
#include <stdint.h>

/* The file is compiled twice into liboracle.so: as is (fp32, the reference's arithmetic type) and through
 * adacof_oracle_f64.c with REAL = double / SUF(x) = x##_f64 -- the high-precision ARBITER used by the
 * "GPU no further from the truth than the reference's own fp32 run" tests. */
#ifndef REAL
#define REAL float
#define SUF(x) x
#endif

static inline int SUF(clampi)(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

/* adacof.py:6-65 */
void SUF(oracle_adacof_forward)(const REAL* input, const REAL* weight, const REAL* off_i,
                           const REAL* off_j, REAL* output, int B, int C, int Hin, int Win,
                           int H, int W, int F, int dil, int i_begin, int i_end) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (int i = i_begin; i < i_end; ++i)      /* row slab: host threads split [0,H) */
            for (int c = 0; c < C; ++c) {
                const REAL* I = input + ((size_t)n * C + c) * plane_in;
                for (int j = 0; j < W; ++j) {
                    REAL acc = (REAL)0;
                    for (int k = 0; k < F; ++k)
                        for (int l = 0; l < F; ++l) {
                            const size_t q = ((size_t)n * F * F + (size_t)k * F + l) * plane + (size_t)i * W + j;
                            const REAL w = weight[q], alpha = off_i[q], beta = off_j[q];
                            const int A = (int)alpha, Bq = (int)beta;          /* :27-28 trunc */
                            const int r0 = SUF(clampi)(i + k * dil + A, Hin - 1);   /* :30-34 */
                            const int c0 = SUF(clampi)(j + l * dil + Bq, Win - 1);  /* :36-40 */
                            const int r1 = SUF(clampi)(i + k * dil + A + 1, Hin - 1);
                            const int c1 = SUF(clampi)(j + l * dil + Bq + 1, Win - 1);
                            const REAL a = alpha - (REAL)A, b = beta - (REAL)Bq;
                            acc += w * (I[(size_t)r0 * Win + c0] * (1 - a) * (1 - b) +
                                        I[(size_t)r1 * Win + c0] * a * (1 - b) +
                                        I[(size_t)r0 * Win + c1] * (1 - a) * b +
                                        I[(size_t)r1 * Win + c1] * a * b);     /* :54-59 */
                        }
                    output[((size_t)n * C + c) * plane + (size_t)i * W + j] = acc;
                }
            }
}

/* adacof.py:67-128 (gW), :130-193 (g_alpha), :195-258 (g_beta); C is 3 in the reference. */
void SUF(oracle_adacof_backward)(const REAL* gout, const REAL* input, const REAL* weight,
                            const REAL* off_i, const REAL* off_j, REAL* gw, REAL* goi,
                            REAL* goj, int B, int C, int Hin, int Win, int H, int W, int F, int dil,
                            int i_begin, int i_end) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (int kl = 0; kl < F * F; ++kl) {
            const int k = kl / F, l = kl % F;
            for (int i = i_begin; i < i_end; ++i)
                for (int j = 0; j < W; ++j) {
                    const size_t q = ((size_t)n * F * F + kl) * plane + (size_t)i * W + j;
                    const REAL w = weight[q], alpha = off_i[q], beta = off_j[q];
                    const int A = (int)alpha, Bq = (int)beta;
                    const int r0 = SUF(clampi)(i + k * dil + A, Hin - 1);
                    const int c0 = SUF(clampi)(j + l * dil + Bq, Win - 1);
                    const int r1 = SUF(clampi)(i + k * dil + A + 1, Hin - 1);
                    const int c1 = SUF(clampi)(j + l * dil + Bq + 1, Win - 1);
                    const REAL a = alpha - (REAL)A, b = beta - (REAL)Bq;
                    REAL sw = (REAL)0, sa = (REAL)0, sb = (REAL)0;
                    for (int c = 0; c < C; ++c) {
                        const REAL* I = input + ((size_t)n * C + c) * plane_in;
                        const REAL d = gout[((size_t)n * C + c) * plane + (size_t)i * W + j];
                        const REAL v00 = I[(size_t)r0 * Win + c0], v10 = I[(size_t)r1 * Win + c0];
                        const REAL v01 = I[(size_t)r0 * Win + c1], v11 = I[(size_t)r1 * Win + c1];
                        sw += d * (v00 * (1 - a) * (1 - b) + v10 * a * (1 - b) + v01 * (1 - a) * b + v11 * a * b); /* :118-123 */
                        sa += d * w * (-v00 * (1 - b) + v10 * (1 - b) - v01 * b + v11 * b);                       /* :183-188 */
                        sb += d * w * (-v00 * (1 - a) - v10 * a + v01 * (1 - a) + v11 * a);                       /* :248-253 */
                    }
                    gw[q] = sw; goi[q] = sa; goj[q] = sb;
                }
        }
}

/* True gradient w.r.t. input (NOT in the reference, which returns zeros --
 * adacof.py:382,445).  Adjoint of the forward above; used only to check the
 * optional gin_mode=2 extension.  Serial scatter, deterministic order. */
void SUF(oracle_adacof_grad_input)(const REAL* gout, const REAL* weight, const REAL* off_i,
                              const REAL* off_j, REAL* gin, int B, int C, int Hin, int Win,
                              int H, int W, int F, int dil) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (size_t t = 0; t < (size_t)B * C * plane_in; ++t) gin[t] = (REAL)0;
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            REAL* G = gin + ((size_t)n * C + c) * plane_in;
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    const REAL d = gout[((size_t)n * C + c) * plane + (size_t)i * W + j];
                    for (int k = 0; k < F; ++k)
                        for (int l = 0; l < F; ++l) {
                            const size_t q = ((size_t)n * F * F + (size_t)k * F + l) * plane + (size_t)i * W + j;
                            const REAL w = weight[q], alpha = off_i[q], beta = off_j[q];
                            const int A = (int)alpha, Bq = (int)beta;
                            const int r0 = SUF(clampi)(i + k * dil + A, Hin - 1);
                            const int c0 = SUF(clampi)(j + l * dil + Bq, Win - 1);
                            const int r1 = SUF(clampi)(i + k * dil + A + 1, Hin - 1);
                            const int c1 = SUF(clampi)(j + l * dil + Bq + 1, Win - 1);
                            const REAL a = alpha - (REAL)A, b = beta - (REAL)Bq;
                            const REAL dw = d * w;
                            G[(size_t)r0 * Win + c0] += dw * (1 - a) * (1 - b);
                            G[(size_t)r1 * Win + c0] += dw * a * (1 - b);
                            G[(size_t)r0 * Win + c1] += dw * (1 - a) * b;
                            G[(size_t)r1 * Win + c1] += dw * a * b;
                        }
                }
        }
}

/* src/fusion_net/fusion_adacofnet.py:198-213 -- occlusion blend and the
 * flow-variance uncertainty mask, restated per pixel. */
void SUF(oracle_adacofnet_tail)(const REAL* t1, const REAL* t2, const REAL* occ,
                           const REAL* w1, const REAL* a1, const REAL* b1,
                           const REAL* w2, const REAL* a2, const REAL* b2,
                           REAL* frame, REAL* mask, int B, int C, int H, int W, int FF,
                           int i_begin, int i_end) {
    const size_t plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (size_t p = (size_t)i_begin * W; p < (size_t)i_end * W; ++p) {
            const REAL o = occ[(size_t)n * plane + p];
            for (int c = 0; c < C; ++c) {
                const size_t q = ((size_t)n * C + c) * plane + p;
                frame[q] = o * t1[q] + (1 - o) * t2[q];                        /* :198 */
            }
            REAL var[2];
            for (int f = 0; f < 2; ++f) {
                const REAL* w = f ? w2 : w1; const REAL* al = f ? a2 : a1; const REAL* be = f ? b2 : b1;
                REAL mi = (REAL)0, mj = (REAL)0;
                for (int t = 0; t < FF; ++t) {                                  /* :204-205 */
                    const size_t q = ((size_t)n * FF + t) * plane + p;
                    mi += w[q] * al[q]; mj += w[q] * be[q];
                }
                REAL vi = (REAL)0, vj = (REAL)0;
                for (int t = 0; t < FF; ++t) {                                  /* :207-208 */
                    const size_t q = ((size_t)n * FF + t) * plane + p;
                    const REAL di = mi - al[q], dj = mj - be[q];
                    vi += w[q] * (di * di); vj += w[q] * (dj * dj);
                }
                var[f] = vi + vj;                                               /* :211 sum(0) */
            }
            REAL m = var[0] > var[1] ? var[0] : var[1];                        /* :211 */
            m = m < (REAL)0 ? (REAL)0 : (m > (REAL)20 ? (REAL)20 : m);                      /* :212 */
            mask[(size_t)n * plane + p] = m / (REAL)20;
        }
}
