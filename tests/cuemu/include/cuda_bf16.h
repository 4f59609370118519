// cuemu shim for the few <cuda_bf16.h> conversions the library uses (tests only): round-to-nearest-even like the device.
#pragma once
#include <stdint.h>
#include <string.h>

struct __nv_bfloat16 { uint16_t bits; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = 0x7fff; return r; }          // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    r.bits = (uint16_t)(u >> 16);
    return r;
}
static inline float __bfloat162float(__nv_bfloat16 h) {
    const uint32_t u = (uint32_t)h.bits << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline unsigned short __bfloat16_as_ushort(__nv_bfloat16 h) { return h.bits; }
static inline __nv_bfloat16 __ushort_as_bfloat16(unsigned short b) { __nv_bfloat16 r; r.bits = b; return r; }

// packed pair (x in the low half), as cvt.rn.bf16x2.f32 produces it
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
