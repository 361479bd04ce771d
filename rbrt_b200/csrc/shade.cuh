// shade.cuh — camera ray generation, the three RayScattering impls and the sky, op for op.
//   cam.rs:64-82, materials.rs:14-37, lambertian.rs:11-24, metal.rs:12-25, dielectric.rs:11-85,
//   lib.rs:68-71.
#pragma once
#include "common.cuh"

namespace rbrt {

struct RngKey { uint32_t k0, k1; };

// ------------------------------------------------------------------ cam.rs:64-82
__device__ __forceinline__ void camera_ray(const CamDev& c, uint32_t row, uint32_t col, RngKey key, uint32_t pixel,
                                           uint32_t sample, f3& o, f3& d) {
    float col_off = XSUB((float)col, (float)(c.width / 2));            // integer halving (cam.rs:65-66)
    float row_off = XSUB((float)row, (float)(c.height / 2));
    u4 r = philox4x32_10(pixel, sample, 0u, 0u, key.k0, key.k1);       // bounce 0, round 0
    float u1 = u32_to_unit_f32(r.x), u2 = u32_to_unit_f32(r.y);        // col draw, then row draw (cam.rs:69,71)
    float col_mm = XMUL(XSUB(XADD(col_off, u1), 0.5f), c.mm_per_pix_hor);
    float row_mm = XMUL(XSUB(XADD(row_off, u2), 0.5f), c.mm_per_pix_vert);
    f3 pos = mk3(c.pos[0], c.pos[1], c.pos[2]);
    f3 right = mk3(c.right[0], c.right[1], c.right[2]);
    f3 up = mk3(c.up[0], c.up[1], c.up[2]);
    f3 center = mk3(c.center[0], c.center[1], c.center[2]);
    f3 target = (center + XMUL(0.001f, col_mm) * right) - XMUL(0.001f, row_mm) * up;
    d = norm3(target - pos);
    o = pos;
}

// ------------------------------------------------------------------ materials.rs:14-30
__device__ __forceinline__ f3 random_point_in_unit_sphere(RngKey key, uint32_t pixel, uint32_t sample, uint32_t bounce) {
    uint32_t round = 0;
    f3 p;
    do {
        u4 r = philox4x32_10(pixel, sample, bounce, round++, key.k0, key.k1);
        p = 2.0f * mk3(u32_to_unit_f32(r.x), u32_to_unit_f32(r.y), u32_to_unit_f32(r.z)) - mk3(1.0f, 1.0f, 1.0f);
    } while (len3(p) > 1.0f);                                           // accepts length == 1 (materials.rs:21)
    return p;
}

// materials.rs:32-37
__device__ __forceinline__ f3 reflect3(f3 d, f3 n) {
    f3 du = norm3(d), nu = norm3(n);
    f3 r = du - (2.0f * nu) * dot3(du, nu);
    return norm3(r);
}

__device__ __forceinline__ float powi2(float x) { return XMUL(x, x); }
__device__ __forceinline__ float powi5(float x) { float x2 = XMUL(x, x); float x4 = XMUL(x2, x2); return XMUL(x, x4); }

// dielectric.rs:63-66
__device__ __forceinline__ float schlick(float cosine, float ref_index) {
    float r0 = powi2(XDIV(XSUB(1.0f, ref_index), XADD(1.0f, ref_index)));
    return XADD(r0, XMUL(XSUB(1.0f, r0), powi5(XSUB(1.0f, cosine))));
}

// dielectric.rs:68-85
__device__ __forceinline__ bool refract3(f3 d, f3 n, float ni_over_nt, f3& out) {
    f3 vu = norm3(d), nu = norm3(n);
    float c = dot3(vu, nu);
    float discr = XSUB(1.0f, XMUL(powi2(ni_over_nt), XSUB(1.0f, powi2(c))));
    if (discr > 0.0f) {
        out = ni_over_nt * (vu - nu * c) - XSQRT(discr) * nu;           // not re-normalised (dielectric.rs:81)
        return true;
    }
    return false;
}

// One scatter.  mat = {albedo.xyz, param}.  Returns false when the path is absorbed (metal.rs:24).
// The attenuation is implied by the element (albedo, or (1,1,1) for dielectrics) and applied when
// the path ends (see radiance unwinding in render.cu).
__device__ __forceinline__ bool scatter(uint32_t kind, float4 mat, f3 in_d, f3 point, f3 normal, RngKey key,
                                        uint32_t pixel, uint32_t sample, uint32_t bounce, f3& out_d) {
    if (kind == 0u) {                                                   // lambertian.rs:11-24
        f3 target = (point + norm3(normal)) + random_point_in_unit_sphere(key, pixel, sample, bounce);
        out_d = norm3(target - point);
        return true;
    } else if (kind == 1u) {                                            // metal.rs:12-25
        f3 refl = reflect3(in_d, normal);
        out_d = norm3(refl + mat.w * random_point_in_unit_sphere(key, pixel, sample, bounce));
        return dot3(out_d, normal) > 0.0f;
    } else {                                                            // dielectric.rs:11-60
        float ref_idx = mat.w;
        f3 refl = reflect3(in_d, normal);
        f3 outward; float ni_over_nt, cosine;
        float a = dot3(norm3(in_d), norm3(normal));
        if (a > 0.0f) { outward = -1.0f * normal; ni_over_nt = ref_idx; cosine = XMUL(ref_idx, a); }
        else { outward = normal; ni_over_nt = XDIV(1.0f, ref_idx); cosine = -a; }
        f3 refr = mk3(0.0f, 0.0f, 0.0f);
        float reflect_prob = refract3(in_d, outward, ni_over_nt, refr) ? schlick(cosine, ref_idx) : 1.0f;
        u4 r = philox4x32_10(pixel, sample, bounce, 0u, key.k0, key.k1);
        out_d = (u32_to_unit_f32(r.x) < reflect_prob) ? refl : refr;
        return true;
    }
}

// lib.rs:68-71 with bg = (0.05, 0.05, 0.8) (lib.rs:89-93)
__device__ __forceinline__ f3 sky(f3 d) {
    float t = XMUL(0.5f, XADD(d.y, 1.0f));
    return t * mk3(1.0f, 1.0f, 1.0f) + XSUB(1.0f, t) * mk3(0.05f, 0.05f, 0.8f);
}

// Rust `as u8`: truncating, saturating, NaN -> 0 (lib.rs:118-120)
__device__ __forceinline__ uint8_t as_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)(int)v;
}

}  // namespace rbrt
