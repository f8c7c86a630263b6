// common.cuh -- shared host/device helpers for libfvfi.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "fvfi.h"

namespace fvfi {

void set_error(const char* fmt, ...);
void note_launch();  // counts kernel launches issued by this library (fvfi_launch_count)

#define FVFI_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ::fvfi::set_error(__VA_ARGS__);  \
            return FVFI_EINVAL;              \
        }                                    \
    } while (0)

#define FVFI_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::fvfi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return FVFI_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

#define FVFI_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        ::fvfi::note_launch();                                                            \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::fvfi::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                        \
            return FVFI_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

constexpr int FVFI_MAX_DEVICES = 64;
int current_device();   // -1 on error
int sm_count();         // SMs of the current device (cached)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) -- raised again only when a launch asks for more --
// instead of a driver call on every launch.  Returns FVFI_OK / FVFI_ECUDA (error text set).
int smem_opt_in(const void* kernel, size_t bytes);
#define FVFI_SMEM_OPT_IN(kernel, bytes)                                             \
    do {                                                                            \
        if (int _rc = ::fvfi::smem_opt_in((const void*)(kernel), (size_t)(bytes))) return _rc; \
    } while (0)

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Streaming (read-once / write-once) global accesses: keep them out of L1 so the gathered
// frame tiles stay resident.
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// Bilinear resampling exactly as ATen's upsample_bilinear2d (area_pixel_compute_source_index): source rows / columns and the
// weight of the second one for output index o.  Shared by the resize kernel (imgproc.cu) and by the convolution loader that
// upsamples its input on the fly (conv_tc.cu), so that both give the same bits.
__host__ __device__ __forceinline__ void bilinear_src(int o, float scale, int align_corners, int n_in, int& i0, int& i1, float& l) {
    const float f = align_corners ? scale * (float)o : fmaxf(scale * ((float)o + 0.5f) - 0.5f, 0.f);
    i0 = min((int)f, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    l = f - (float)i0;
}
__host__ __device__ __forceinline__ float bilinear_scale(int n_in, int n_out, int align_corners) {
    if (align_corners) return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
    return (float)n_in / (float)n_out;
}
__device__ __forceinline__ float bilerp(float hy, float hx, float ly, float lx, float a, float b, float c, float d) {
    return fmaf(ly, fmaf(lx, d, hx * c), hy * fmaf(lx, b, hx * a));
}
}  // namespace fvfi
