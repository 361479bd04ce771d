// cuda_runtime.h — NOT the CUDA header: a stand-in found first on the include path when tests/host_device/harness.cpp compiles
// the product's device headers (rbrt_b200/csrc/common.cuh, intersect.cuh, shade.cuh) with g++ for the HOST.  TEST INFRASTRUCTURE:
// it lets `pytest -m "not gpu"` run the source of the product's per-ray arithmetic against the oracle in a container without a
// GPU.  Nothing under rbrt_b200/ includes it; the product has no CPU path.
// It provides only what those three headers touch: the qualifiers, float4 / uint4, and a handful of intrinsics with their IEEE meaning.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }          // (only in the BVH slab tests, which the host build never runs)
// The device's __fdividef(1, x) is MUFU.RCP, good to about 1 ulp, not the exact quotient: hd_rcp_error (set through hd_set_rcp_error) scales the
// three reciprocals of ray_slabs by (1 + e), (1 - e), (1 + e) in turn — the pattern that moves the slabs of different axes apart — so the tests can
// show that the traversal's margins also cover the device's reciprocal.
extern float hd_rcp_error;
extern unsigned hd_rcp_calls;
static inline float __fdividef(float a, float b) { const float q = a / b; return hd_rcp_error == 0.0f ? q : q * (1.0f + ((hd_rcp_calls++ % 3u) == 1u ? -hd_rcp_error : hd_rcp_error)); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }                            // (the device's is an approximation: the cull tests perturb its result)
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __ballot_sync(uint32_t, int pred) { return pred ? 1u : 0u; }          // a "warp" of one lane
