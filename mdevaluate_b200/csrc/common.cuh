// Shared helpers for the mdevaluate_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mdevaluate_b200.h"

namespace mdb {

// ---- error reporting across the C ABI (thread-local text, never throws) --------------------
char *error_buffer();
int fail(int status, const char *fmt, ...);
int &last_launch_count();

#define MDB_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            return ::mdb::fail(MDB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                               cudaGetErrorString(err__), __FILE__, __LINE__);                \
    } while (0)

// ---- small device utilities ----------------------------------------------------------------

// int -> double on the FP64 pipe instead of the conversion (XU) pipe: 2^52 + 2^31 + n, minus the bias.
__device__ __forceinline__ double int_to_double(int n) {
    return __hiloint2double(0x43300000, n ^ 0x80000000) - 4503601774854144.0;
}

// Image counts are stored excess-128 (uint8 = n + 128): byte j of a packed word becomes the double n
// with one byte-permute and one DADD on the FP64 pipe (2^52 + b, minus 2^52 + 128) -- no sign
// extension, no int->double conversion on the (contended) conversion pipe.
__device__ __forceinline__ double image_count(uint32_t w, int j) {
    const uint32_t b = (w >> (8 * j)) & 0xffu;
    return __hiloint2double(0x43300000, (int)b) - 4503599627370624.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load: data is read once per (origin, lag) -> do not keep it in L1
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
    double2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// sin(sqrt(y)) / sqrt(y) for y >= 0 in float32, |error| ~1e-7 for sqrt(y) up to a few 1e4.
//   one MUFU.RSQ + a Newton step give 1/x and x = y/x;
//   x is reduced by multiples of pi (2-term Cody-Waite with FMA), sin on [-pi/2, pi/2] is an odd
//   degree-11 polynomial (near-minimax, 1e-10), the sign comes from the parity of the multiple.
__device__ __forceinline__ float sinc_of_square(float y) {
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(y));   // bare MUFU.RSQ; y < 1e-30 is handled below
    rs = rs * fmaf(-0.5f * y * rs, rs, 1.5f);
    const float x = y * rs;
    const float kf_biased = fmaf(x, 0.318309886183790672f, 12582912.0f);
    const uint32_t sign = ((uint32_t)__float_as_int(kf_biased)) << 31;
    const float kf = kf_biased - 12582912.0f;
    float z = fmaf(-kf, 3.140625f, x);
    z = fmaf(-kf, 9.676535846665502e-4f, z);   // pi - 3.140625 - this = 5e-12: below float32 for k < 1e4
    const float s = z * z;
    float p = -2.4061128200969506e-08f;
    p = fmaf(p, s, 2.7535631943464978e-06f);
    p = fmaf(p, s, -1.9841075118165463e-04f);
    p = fmaf(p, s, 8.33333283662796e-03f);
    p = fmaf(p, s, -1.666666716337204e-01f);
    float sn = fmaf(z * s, p, z);
    sn = __int_as_float(__float_as_int(sn) ^ sign);
    const float v = sn * rs;
    return (y < 1e-30f) ? 1.0f : v;
}

}  // namespace mdb
