// cull.cuh — the cone tests behind the camera-ray culling of stage A (render.cu): in a header of their own so that
// tests/host_device can compile them for the host and check, ray by ray, that they only ever skip a test the exact arithmetic
// (sphere_intersect / mesh_bbox_hit, intersect.cuh) answers with "no hit".  Included by render.cu where the functions used to stand.
#pragma once
#include "common.cuh"

namespace rbrt {

__device__ __forceinline__ float4 cone_of_sphere(f3 o, float cx, float cy, float cz, float r) {
    const float lx = cx - o.x, ly = cy - o.y, lz = cz - o.z;
    const float d2 = lx * lx + ly * ly + lz * lz, r2 = r * r;
    // The reference's own arithmetic works on absolute coordinates: o - c and (bound - o) carry an error of ~ulp(|coordinate|), which a cone
    // margin relative to the DISTANCE only covers while the distance is not tiny against the coordinates (camera or element far from the
    // world origin): there, always test.
    const float mag2 = (o.x * o.x + o.y * o.y + o.z * o.z) + (cx * cx + cy * cy + cz * cz);
    if (!(d2 > 1.01f * r2) || !(d2 > 1e-20f) || !(d2 < 1e30f) || !(d2 > 1e-4f * mag2)) return make_float4(0.0f, 0.0f, 0.0f, -2.0f);   // inside / too close / degenerate: always test
    const float inv = rsqrtf(d2);
    return make_float4(lx * inv, ly * inv, lz * inv, (1.0f - r2 / d2) - 1e-4f);
}
// d is a camera ray's direction: unit length to ~1e-7 (cam.rs:80 normalises it), which the 1e-4 margin absorbs.
// outside_cone (mesh boxes): u|u| < cos^2 - margin covers both "points away" (u < 0: the slab test rejects t_max < 0, aabbox.rs:49) and
// "outside the cone".  outside_double_cone (spheres): u^2 < cos^2 - margin, i.e. the whole LINE misses the sphere — a sphere BEHIND the
// origin is not a provable miss: with a discriminant of exactly 0 the reference keeps the single root even when it is negative
// (sphere.rs:34-44: the far root is only tried when there are two) and reports a hit behind the ray.  Found by the round-2 parity runs
// (an experiment with sphere groups for bounce rays skipped such a sphere: 1 path in 6.3 M on C4 differed), pinned in
// tests/test_oracle_quirks.py (Q16) and tests/test_gpu_trace.py.  q.w = -2 ("always test") and NaN directions compare false -> tested.
__device__ __forceinline__ bool outside_cone(float4 q, f3 d) {
    const float u = __fmaf_rn(d.z, q.z, __fmaf_rn(d.y, q.y, d.x * q.x));
    return u * fabsf(u) < q.w;
}
__device__ __forceinline__ bool outside_double_cone(float4 q, f3 d) {
    const float u = __fmaf_rn(d.z, q.z, __fmaf_rn(d.y, q.y, d.x * q.x));
    return u * u < q.w;
}

}  // namespace rbrt
