// fft_plan.hpp -- host-side planning for fft_engine.cuh: radix selection, Bluestein length, tables (all in double).
#pragma once
#include <math.h>

#include <algorithm>
#include <complex>
#include <vector>

#include "fft_engine.cuh"

namespace fvfi {

struct HostFftPlan {
    FftPlan p{};                       // pointers are filled in by whoever owns the storage (device upload or host test)
    std::vector<float2> tw, chirp, bhat;
    std::vector<unsigned short> perm;
};

// Split n into the radices the engine implements; false if n has a prime factor > 19.
inline bool fft_radices(int n, std::vector<int>& out) {
    out.clear();
    int e2 = 0, e3 = 0, e5 = 0;
    std::vector<int> big;
    int rem = n;
    while (rem % 2 == 0) { rem /= 2; ++e2; }
    while (rem % 3 == 0) { rem /= 3; ++e3; }
    while (rem % 5 == 0) { rem /= 5; ++e5; }
    for (int p : {7, 11, 13, 17, 19})
        while (rem % p == 0) { rem /= p; big.push_back(p); }
    if (rem != 1) return false;
    std::vector<int> odd(big);
    while (e3 >= 1 && e5 >= 1 && (e3 & 1)) { odd.push_back(15); --e3; --e5; }
    while (e3 >= 2) { odd.push_back(9); e3 -= 2; }
    // a lone 2 or 4 would cost a full pass: merge it with a lone 3 / 5 (radix 6, 10, 12)
    std::vector<int> pow2;
    while (e2 >= 4 && e2 != 5 && e2 != 6) { pow2.push_back(16); e2 -= 4; }
    if (e2 == 6) { pow2.push_back(8); pow2.push_back(8); e2 = 0; }
    if (e2 == 5) { pow2.push_back(8); pow2.push_back(4); e2 = 0; }
    if (e2 == 3) { pow2.push_back(8); e2 = 0; }
    if (e2 == 2) {
        if (e3 == 1) { odd.push_back(12); e3 = 0; } else pow2.push_back(4);
        e2 = 0;
    }
    if (e2 == 1) {
        if (e3 == 1) { odd.push_back(6); e3 = 0; }
        else if (e5 >= 1) { odd.push_back(10); --e5; }
        else pow2.push_back(2);
        e2 = 0;
    }
    while (e3 >= 1) { odd.push_back(3); --e3; }
    while (e5 >= 1) { odd.push_back(5); --e5; }
    std::sort(pow2.begin(), pow2.end(), [](int a, int b) { return a > b; });
    std::sort(odd.begin(), odd.end(), [](int a, int b) { return (a & 1) != (b & 1) ? (a & 1) < (b & 1) : a > b; });   // 12/10/6 before odd
    out = pow2;                        // DIF order: power-of-two radices first, odd (conflict-free strides) last
    out.insert(out.end(), odd.begin(), odd.end());
    if (out.empty()) out.push_back(1);
    return (int)out.size() <= FFT_MAX_STAGES;
}

// Reference DFT in double (mixed radix recursion, naive at prime leaves) for the plan tables.
inline void fft_host_dft(std::vector<std::complex<double>>& x) {
    const int n = (int)x.size();
    if (n <= 1) return;
    int r = n;
    for (int d = 2; d * d <= n; ++d)
        if (n % d == 0) { r = d; break; }
    if (r == n) {
        std::vector<std::complex<double>> y(n);
        for (int k = 0; k < n; ++k) {
            std::complex<double> s = 0;
            for (int j = 0; j < n; ++j) s += x[j] * std::polar(1.0, -2.0 * M_PI * (double)((long long)j * k % n) / n);
            y[k] = s;
        }
        x = y;
        return;
    }
    const int m = n / r;
    std::vector<std::vector<std::complex<double>>> sub(r, std::vector<std::complex<double>>(m));
    for (int j = 0; j < n; ++j) sub[j % r][j / r] = x[j];
    for (int q = 0; q < r; ++q) fft_host_dft(sub[q]);
    for (int k = 0; k < n; ++k) {
        std::complex<double> s = 0;
        for (int q = 0; q < r; ++q) s += sub[q][k % m] * std::polar(1.0, -2.0 * M_PI * (double)((long long)q * k % n) / n);
        x[k] = s;
    }
}

inline unsigned fft_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + d - 1) / d); }

// Build the plan for length n.  stockham = true: out-of-place autosort order for direct lengths (rows).
inline bool fft_make_plan(int n, bool stockham, HostFftPlan& H) {
    FftPlan& p = H.p;
    p = FftPlan{};
    p.n = n;
    std::vector<int> rad;
    if (fft_radices(n, rad)) {
        p.M = n;
        p.bluestein = 0;
    } else {
        // Bluestein: smooth M >= 2n-1 minimising M * stages
        long best = -1;
        int bestM = 0;
        for (int M = 2 * n - 1; M <= 4 * n; ++M) {
            std::vector<int> r;
            int rem = M;
            for (int q : {2, 3, 5})
                while (rem % q == 0) rem /= q;
            if (rem != 1 || !fft_radices(M, r)) continue;
            const long cost = (long)M * (long)r.size();
            if (best < 0 || cost < best) { best = cost; bestM = M; }
        }
        if (best < 0) return false;
        p.M = bestM;
        p.bluestein = 1;
        fft_radices(bestM, rad);
        stockham = false;
    }
    const int M = p.M;
    if (M > 65535) return false;
    if (stockham) std::reverse(rad.begin(), rad.end());   // odd radix first: its stride-R stores are conflict free
    p.nfac = (int)rad.size();
    if (rad.size() == 1 && rad[0] == 1) p.nfac = 0;
    int prod = 1;
    for (int s = 0; s < p.nfac; ++s) {
        p.fac[s] = rad[s];
        if (stockham) {
            p.sub[s] = prod;                 // Ns
            prod *= rad[s];
        } else {
            prod *= rad[s];
            p.sub[s] = M / prod;             // m_s
        }
        p.mag_sub[s] = fft_magic((unsigned)p.sub[s]);
        p.mag_items[s] = fft_magic((unsigned)(M / rad[s]));
    }
    // smallest butterfly stride: first Stockham radix / last DIF radix; even -> skewed rows
    p.pad = (p.nfac > 0 && ((stockham ? p.fac[0] : p.fac[p.nfac - 1]) % 2 == 0)) ? 1 : 0;
    H.tw.resize(M);
    for (int t = 0; t < M; ++t) {
        const double a = -2.0 * M_PI * (double)t / (double)M;
        H.tw[t] = make_float2((float)cos(a), (float)sin(a));
    }
    // digit reversal of the DIF network: position p = sum d_s * m_s  <->  natural k = sum d_s * prod_{i<s} fac_i
    H.perm.assign(M, 0);
    if (!stockham) {
        for (int pos = 0; pos < M; ++pos) {
            int rem = pos, k = 0, w = 1;
            for (int s = 0; s < p.nfac; ++s) {
                const int d = rem / p.sub[s];
                rem -= d * p.sub[s];
                k += d * w;
                w *= p.fac[s];
            }
            H.perm[pos] = (unsigned short)k;
        }
    }
    H.chirp.clear();
    H.bhat.clear();
    if (p.bluestein) {
        H.chirp.resize(n);
        std::vector<std::complex<double>> b(M, 0.0);
        for (int k = 0; k < n; ++k) {
            const long long k2 = ((long long)k * k) % (2LL * n);     // exact angle reduction
            const double a = M_PI * (double)k2 / (double)n;
            H.chirp[k] = make_float2((float)cos(a), (float)-sin(a));   // exp(-i pi k^2 / n)
            const std::complex<double> bk = std::polar(1.0, a);        // exp(+i pi k^2 / n)
            b[k] = bk;
            if (k) b[M - k] = bk;
        }
        fft_host_dft(b);
        H.bhat.resize(M);
        for (int pos = 0; pos < M; ++pos) {
            const std::complex<double> v = b[H.perm[pos]] / (double)M;
            H.bhat[pos] = make_float2((float)v.real(), (float)v.imag());
        }
    }
    return true;
}

}  // namespace fvfi
