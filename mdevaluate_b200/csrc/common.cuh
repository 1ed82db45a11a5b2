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

// ---- packed float32 pairs: sm_100 executes add/mul/fma on two floats per instruction (FFMA2) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 bcast2(float c) { return pack2(c, c); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// Byte j of a packed excess-128 image word as the float 2^23 + byte: ONE byte-permute splices the byte
// into the mantissa of 8388608.0f.  Differences of two such floats are the (exact) differences of the
// image counts -- no integer-to-float conversion.
__device__ __forceinline__ float image_float(uint32_t w, int j) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)j));
}
// The same with the constant 0x4B000000 handed in as a run-time REGISTER value (a kernel parameter): PRMT takes one
// immediate, and with the constant in a register it is the selector -- otherwise ptxas materialises twelve selectors
// in registers per row (measured in the SASS: 12 extra moves per 4 evaluations).
__device__ __forceinline__ float image_float(uint32_t w, int j, uint32_t magic) {
    return __uint_as_float(__byte_perm(w, magic, 0x7650u | (uint32_t)j));
}
__device__ __forceinline__ float rsqrt_approx(float y) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));   // bare MUFU.RSQ
    return r;
}

// sin(sqrt(y)) / sqrt(y) for two values y >= 0 at once, float32, |error| ~1e-7 for sqrt(y) up to a
// few 1e4:  one MUFU.RSQ + a Newton step give 1/x and x = y/x;  x is reduced by multiples of pi
// (2-term Cody-Waite with FMA), sin on [-pi/2, pi/2] is an odd degree-11 polynomial (near-minimax,
// 1e-10), the sign comes from the parity of the multiple.  Every float32 add / mul / fma is one
// packed instruction for the pair.  y is offset by 2^-100 (a power of four, so rsqrt, the Newton
// step and the final product are exact there): y = 0 gives exactly 1.  Valid for sqrt(y) = q*r < 2^22*pi
// (1.3e7 rad; beyond that the nearest-multiple trick runs out of mantissa) -- MD displacements times any
// physical q stay orders of magnitude below.
__device__ __forceinline__ void sinc_of_square_x2(float y0, float y1, float &r0, float &r1) {
    // y + 2^-100 instead of max(y, 2^-100): one packed add; any y that matters (> 2^-76) is unchanged by it
    const f32x2 y = add2(pack2(y0, y1), bcast2(7.888609052210118e-31f));
    unpack2(y, y0, y1);
    f32x2 rs = pack2(rsqrt_approx(y0), rsqrt_approx(y1));
    f32x2 t = mul2(mul2(y, bcast2(-0.5f)), rs);
    t = fma2(t, rs, bcast2(1.5f));
    rs = mul2(rs, t);
    const f32x2 x = mul2(y, rs);
    const f32x2 kb = fma2(x, bcast2(0.318309886183790672f), bcast2(12582912.0f));
    const f32x2 kf = add2(kb, bcast2(-12582912.0f));
    f32x2 z = fma2(kf, bcast2(-3.140625f), x);
    z = fma2(kf, bcast2(-9.676535846665502e-4f), z);
    const f32x2 s = mul2(z, z);
    f32x2 p = fma2(s, bcast2(-2.4061128200969506e-08f), bcast2(2.7535631943464978e-06f));
    p = fma2(p, s, bcast2(-1.9841075118165463e-04f));
    p = fma2(p, s, bcast2(8.33333283662796e-03f));
    p = fma2(p, s, bcast2(-1.666666716337204e-01f));
    f32x2 sn = fma2(mul2(z, s), p, z);
    float kb0, kb1, sn0, sn1;
    unpack2(kb, kb0, kb1);
    unpack2(sn, sn0, sn1);
    // parity of the multiple -> sign: adding (kb << 31) flips bit 31 exactly like the xor would, and the
    // shift-and-add is ONE integer multiply-add
    sn0 = __uint_as_float(__float_as_uint(sn0) + (__float_as_uint(kb0) << 31));
    sn1 = __uint_as_float(__float_as_uint(sn1) + (__float_as_uint(kb1) << 31));
    unpack2(mul2(pack2(sn0, sn1), rs), r0, r1);
}

// ---- image-count formats ------------------------------------------------------------------------------
// IMG_BYTES:  one byte per component, excess-128 (the scan's native output);
// IMG_PACKED: one 32-bit word per group of 4 atoms, byte a = atom a of the group, bits 0-1 / 2-3 / 4-5 =
//             n_x + 1, n_y + 1, n_z + 1 (|n| <= 1, checked by the pack kernel); 1 B instead of 3 B per atom;
// IMG_I16:    int16 per component (|n| > 127; correctness fallback, direct-load kernel only).
enum { IMG_NONE = 0, IMG_BYTES = 1, IMG_PACKED = 2, IMG_I16 = 3 };

// Masking a packed word with 0x03030303 << 2d leaves the fields of dimension d of the four atoms in separate
// bytes, scaled by 4^d.  The scale is a power of two, so it is folded into the box (box_d / 4^d): products and
// FMA results are bit-identical to the unscaled ones.
__device__ __forceinline__ uint32_t packed_dim(uint32_t w, int d) { return w & (0x03030303u << (2 * d)); }
// byte a of a masked word as the double 4^d * n (n = field - 1)
__device__ __forceinline__ double packed_count(uint32_t masked, int a, int d) {
    const uint32_t b = __byte_perm(masked, 0u, 0x4440u | (uint32_t)a);
    return __hiloint2double(0x43300000, (int)b) - (4503599627370496.0 + (double)(1 << (2 * d)));
}

// predicated shared-memory increment (no return value): red.shared stays a single predicated instruction
// Unconditional shared-memory add without return value.  ptxas wraps a PREDICATED red/atom in a branch with its own
// convergence region (BSSY / BRA / BSYNC per update, seen in the SASS), so the histogram kernels instead redirect the
// samples that must not count to a scratch word; `value` comes from a register (a kernel parameter): an immediate 1
// turns the instruction into the warp-aggregating ATOMS.POPC.INC, which again needs the convergence region.
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t value) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(value) : "memory");
}

// ---- mbarrier + bulk-copy (TMA engine) primitives shared by the streaming kernels ------------------
__device__ __forceinline__ uint32_t smem_u32(const void *ptr) {
    return (uint32_t)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive executed by the lanes with pred != 0 (predicated, no divergent branch)
__device__ __forceinline__ void mbar_arrive_if(uint32_t bar, uint32_t pred) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.u32 P1, %1, 0;\n"
        "@P1 mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}" ::"r"(bar),
        "r"(pred)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy executed by the TMA engine; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
}  // namespace mdb
