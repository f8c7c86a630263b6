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

static inline int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

/* adacof.py:6-65 */
void oracle_adacof_forward(const float* input, const float* weight, const float* off_i,
                           const float* off_j, float* output, int B, int C, int Hin, int Win,
                           int H, int W, int F, int dil, int i_begin, int i_end) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (int i = i_begin; i < i_end; ++i)      /* row slab: host threads split [0,H) */
            for (int c = 0; c < C; ++c) {
                const float* I = input + ((size_t)n * C + c) * plane_in;
                for (int j = 0; j < W; ++j) {
                    float acc = 0.0f;
                    for (int k = 0; k < F; ++k)
                        for (int l = 0; l < F; ++l) {
                            const size_t q = ((size_t)n * F * F + (size_t)k * F + l) * plane + (size_t)i * W + j;
                            const float w = weight[q], alpha = off_i[q], beta = off_j[q];
                            const int A = (int)alpha, Bq = (int)beta;          /* :27-28 trunc */
                            const int r0 = clampi(i + k * dil + A, Hin - 1);   /* :30-34 */
                            const int c0 = clampi(j + l * dil + Bq, Win - 1);  /* :36-40 */
                            const int r1 = clampi(i + k * dil + A + 1, Hin - 1);
                            const int c1 = clampi(j + l * dil + Bq + 1, Win - 1);
                            const float a = alpha - (float)A, b = beta - (float)Bq;
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
void oracle_adacof_backward(const float* gout, const float* input, const float* weight,
                            const float* off_i, const float* off_j, float* gw, float* goi,
                            float* goj, int B, int C, int Hin, int Win, int H, int W, int F, int dil,
                            int i_begin, int i_end) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (int kl = 0; kl < F * F; ++kl) {
            const int k = kl / F, l = kl % F;
            for (int i = i_begin; i < i_end; ++i)
                for (int j = 0; j < W; ++j) {
                    const size_t q = ((size_t)n * F * F + kl) * plane + (size_t)i * W + j;
                    const float w = weight[q], alpha = off_i[q], beta = off_j[q];
                    const int A = (int)alpha, Bq = (int)beta;
                    const int r0 = clampi(i + k * dil + A, Hin - 1);
                    const int c0 = clampi(j + l * dil + Bq, Win - 1);
                    const int r1 = clampi(i + k * dil + A + 1, Hin - 1);
                    const int c1 = clampi(j + l * dil + Bq + 1, Win - 1);
                    const float a = alpha - (float)A, b = beta - (float)Bq;
                    float sw = 0.0f, sa = 0.0f, sb = 0.0f;
                    for (int c = 0; c < C; ++c) {
                        const float* I = input + ((size_t)n * C + c) * plane_in;
                        const float d = gout[((size_t)n * C + c) * plane + (size_t)i * W + j];
                        const float v00 = I[(size_t)r0 * Win + c0], v10 = I[(size_t)r1 * Win + c0];
                        const float v01 = I[(size_t)r0 * Win + c1], v11 = I[(size_t)r1 * Win + c1];
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
void oracle_adacof_grad_input(const float* gout, const float* weight, const float* off_i,
                              const float* off_j, float* gin, int B, int C, int Hin, int Win,
                              int H, int W, int F, int dil) {
    const size_t plane_in = (size_t)Hin * Win, plane = (size_t)H * W;
    for (size_t t = 0; t < (size_t)B * C * plane_in; ++t) gin[t] = 0.0f;
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            float* G = gin + ((size_t)n * C + c) * plane_in;
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    const float d = gout[((size_t)n * C + c) * plane + (size_t)i * W + j];
                    for (int k = 0; k < F; ++k)
                        for (int l = 0; l < F; ++l) {
                            const size_t q = ((size_t)n * F * F + (size_t)k * F + l) * plane + (size_t)i * W + j;
                            const float w = weight[q], alpha = off_i[q], beta = off_j[q];
                            const int A = (int)alpha, Bq = (int)beta;
                            const int r0 = clampi(i + k * dil + A, Hin - 1);
                            const int c0 = clampi(j + l * dil + Bq, Win - 1);
                            const int r1 = clampi(i + k * dil + A + 1, Hin - 1);
                            const int c1 = clampi(j + l * dil + Bq + 1, Win - 1);
                            const float a = alpha - (float)A, b = beta - (float)Bq;
                            const float dw = d * w;
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
void oracle_adacofnet_tail(const float* t1, const float* t2, const float* occ,
                           const float* w1, const float* a1, const float* b1,
                           const float* w2, const float* a2, const float* b2,
                           float* frame, float* mask, int B, int C, int H, int W, int FF,
                           int i_begin, int i_end) {
    const size_t plane = (size_t)H * W;
    for (int n = 0; n < B; ++n)
        for (size_t p = (size_t)i_begin * W; p < (size_t)i_end * W; ++p) {
            const float o = occ[(size_t)n * plane + p];
            for (int c = 0; c < C; ++c) {
                const size_t q = ((size_t)n * C + c) * plane + p;
                frame[q] = o * t1[q] + (1 - o) * t2[q];                        /* :198 */
            }
            float var[2];
            for (int f = 0; f < 2; ++f) {
                const float* w = f ? w2 : w1; const float* al = f ? a2 : a1; const float* be = f ? b2 : b1;
                float mi = 0.0f, mj = 0.0f;
                for (int t = 0; t < FF; ++t) {                                  /* :204-205 */
                    const size_t q = ((size_t)n * FF + t) * plane + p;
                    mi += w[q] * al[q]; mj += w[q] * be[q];
                }
                float vi = 0.0f, vj = 0.0f;
                for (int t = 0; t < FF; ++t) {                                  /* :207-208 */
                    const size_t q = ((size_t)n * FF + t) * plane + p;
                    const float di = mi - al[q], dj = mj - be[q];
                    vi += w[q] * (di * di); vj += w[q] * (dj * dj);
                }
                var[f] = vi + vj;                                               /* :211 sum(0) */
            }
            float m = var[0] > var[1] ? var[0] : var[1];                        /* :211 */
            m = m < 0.0f ? 0.0f : (m > 20.0f ? 20.0f : m);                      /* :212 */
            mask[(size_t)n * plane + p] = m / 20.0f;
        }
}
