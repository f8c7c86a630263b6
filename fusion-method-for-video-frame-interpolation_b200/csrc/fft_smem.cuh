// fft_smem.cuh -- shared-memory Stockham FFT of ARBITRARY length for sm_100a.
//
// The sqrt(2)-scale steerable pyramid needs 1-D transforms of every length the level-size rule
// produces (1080p: 1080,764=4*191,540,382,...  1920,1358=2*7*97,679,241,...), so the transform
// is a mixed-radix autosort (Stockham) FFT held entirely in shared memory: hard-wired radix
// 2/3/4/5 butterflies, and a generic radix-r stage (any prime r, O(n*r)) for the rest.
// Twiddles come from a per-length table W_n[k] = exp(-2*pi*i*k/n) computed in double on the
// host (plan), so accuracy does not depend on fast-math sincos.
//
// Data layout: `batch` independent sequences, element (b, i) at buf[b*sb + i*si]
//   rows pass   : si = 1,  sb = n      (batch of rows)
//   column pass : si = CT, sb = 1      (tile of CT adjacent columns)
#pragma once
#include <cuda_runtime.h>

namespace fvfi {

constexpr int FFT_MAX_FACTORS = 16;

struct Fft1D {
    int n;
    int nfac;
    int fac[FFT_MAX_FACTORS];
    const float2* tw;  // device table, n entries
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int idx) {
    float2 w = __ldg(tw + idx);
    if (INV) w.y = -w.y;
    return w;
}

template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2* v) {
    if (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if (R == 3) {
        const float2 t1 = cadd(v[1], v[2]);
        const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 d = csub(v[1], v[2]);
        const float2 r = rot90<INV>(make_float2(0.86602540378443865f * d.x, 0.86602540378443865f * d.y));
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, r);
        v[2] = csub(t2, r);
    } else if (R == 4) {
        const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        const float2 c = cadd(v[1], v[3]), d = rot90<INV>(csub(v[1], v[3]));
        v[0] = cadd(a, c);
        v[1] = cadd(b, d);
        v[2] = csub(a, c);
        v[3] = csub(b, d);
    } else if (R == 5) {
        constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
        constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
        const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
        const float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        const float2 p1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        const float2 p2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        const float2 q1 = rot90<INV>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
        const float2 q2 = rot90<INV>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
        v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = cadd(p1, q1);
        v[4] = csub(p1, q1);
        v[2] = cadd(p2, q2);
        v[3] = csub(p2, q2);
    }
}

// One Stockham stage with a hard-wired radix.  Ns = product of the radices already applied.
template <int R, bool INV>
__device__ __forceinline__ void stage_small(const Fft1D& P, const float2* __restrict__ in, float2* __restrict__ out,
                                            int Ns, int batch, int sb, int si) {
    const int n = P.n, m = n / R, step = n / (Ns * R);
    const int work = batch * m;
    for (int q = threadIdx.x; q < work; q += blockDim.x) {
        int b, j;
        if (si == 1) { b = q / m; j = q - b * m; } else { j = q / batch; b = q - j * batch; }
        const int k = j % Ns;
        float2 v[R];
        const float2* src = in + b * sb + j * si;
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = src[u * m * si];
        if (Ns > 1) {
#pragma unroll
            for (int u = 1; u < R; ++u) v[u] = cmul(v[u], twiddle<INV>(P.tw, u * k * step));
        }
        dft_small<R, INV>(v);
        float2* dst = out + b * sb + ((j - k) * R + k) * si;
#pragma unroll
        for (int u = 0; u < R; ++u) dst[u * Ns * si] = v[u];
    }
}

// Generic radix (any r): one work item per OUTPUT element, r complex MACs each.
template <bool INV>
__device__ __forceinline__ void stage_generic(const Fft1D& P, int r, const float2* __restrict__ in,
                                              float2* __restrict__ out, int Ns, int batch, int sb, int si) {
    const int n = P.n, m = n / r, step = n / (Ns * r);
    const int work = batch * n;
    for (int q = threadIdx.x; q < work; q += blockDim.x) {
        int b, o;
        if (si == 1) { b = q / n; o = q - b * n; } else { o = q / batch; b = q - o * batch; }
        const int k = o % Ns, t1 = o / Ns;
        const int v = t1 % r, j = (t1 / r) * Ns + k;
        const float2* src = in + b * sb + j * si;
        float2 acc = make_float2(0.f, 0.f);
        int e = 0, t = 0;  // e = u*k*step (< n), t = (u*v) mod r
        const int ks = k * step;
        for (int u = 0; u < r; ++u) {
            int idx = e + t * m;
            if (idx >= n) idx -= n;
            const float2 w = twiddle<INV>(P.tw, idx);
            const float2 x = src[u * m * si];
            acc.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc.x));
            acc.y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc.y));
            e += ks;
            t += v;
            if (t >= r) t -= r;
        }
        out[b * sb + o * si] = acc;
    }
}

// In-smem FFT of `batch` sequences.  `a` holds the input; `b` is scratch of the same size.
// All threads of the CTA must call this; the data must be visible (caller syncs before).
// Returns the buffer that holds the result (a or b); a trailing __syncthreads() is included.
template <bool INV>
__device__ __forceinline__ float2* fft_smem(const Fft1D& P, float2* a, float2* b, int batch, int sb, int si) {
    int Ns = 1;
    for (int s = 0; s < P.nfac; ++s) {
        const int r = P.fac[s];
        switch (r) {
            case 2: stage_small<2, INV>(P, a, b, Ns, batch, sb, si); break;
            case 3: stage_small<3, INV>(P, a, b, Ns, batch, sb, si); break;
            case 4: stage_small<4, INV>(P, a, b, Ns, batch, sb, si); break;
            case 5: stage_small<5, INV>(P, a, b, Ns, batch, sb, si); break;
            default: stage_generic<INV>(P, r, a, b, Ns, batch, sb, si); break;
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= r;
    }
    return a;
}

}  // namespace fvfi
