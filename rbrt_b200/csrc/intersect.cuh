// intersect.cuh — closest-hit query = Scene::hit (rbrt_lib/src/scene.rs:19-43) on the GPU.
//
// The reference brute-forces every triangle of a mesh for every ray that passes the mesh AABB
// (mesh.rs:233-243 -> triangle.rs:134-262 -> triangle.rs:392-410).  Here a per-mesh BVH only
// skips triangles that cannot win; the per-triangle arithmetic, the acceptance window, the
// first-index tie rule and the per-mesh / cross-element selection are the reference's, op for op.
#pragma once
#include "common.cuh"

namespace rbrt {

#define RBRT_MIN_DIST 0.001f       // lib.rs:44
#define RBRT_MAX_DIST 2000.0f      // lib.rs:45
#define RBRT_T_CAP 999.99994f      // 1.0f / 0.001f as f32 (triangle.rs:146): triangle t must be < this
#define RBRT_STACK 96

struct Hit {
    int kind;            // -1 none, 0 sphere, 1 mesh, -2 NaN (reference panics, sphere.rs:33)
    uint32_t elem;       // sphere index or mesh index
    uint32_t tri;        // original triangle index
    float t, dist;
};

struct TraceCounters { uint32_t nodes, tris; };

// ------------------------------------------------------------------ sphere.rs:20-66
// returns 1 hit, 0 miss, -1 NaN discriminant
__device__ __forceinline__ int sphere_intersect(float4 s, f3 o, f3 d, float& t_out, float& dist_out) {
    f3 c = mk3(s.x, s.y, s.z);
    float a = dot3(d, d);
    f3 l = o - c;
    float b = dot3(d * 2.0f, l);
    float cc = XSUB(dot3(l, l), XMUL(s.w, s.w));
    float sol = XSUB(XMUL(b, b), XMUL(XMUL(4.0f, a), cc));
    if (sol != sol) return -1;
    if (sol < 0.0f) return 0;
    float sq = XSQRT(sol);
    float two_a = XMUL(2.0f, a);
    float t = XDIV(XSUB(-b, sq), two_a);
    if (sol > 0.0f && t < 0.0f) {
        t = XDIV(XADD(-b, sq), two_a);
        if (t < 0.0f) return 0;
    }
    f3 p = o + t * d;
    float dist = len3(o - p);
    if (dist < RBRT_MIN_DIST || dist > RBRT_MAX_DIST) return 0;
    t_out = t; dist_out = dist;
    return 1;
}

// ------------------------------------------------------------------ aabbox.rs:28-58 (whole-mesh pre-test)
__device__ __forceinline__ bool mesh_bbox_hit(const MeshDev& m, f3 o, f3 d) {
    float tlx = XDIV(XSUB(m.lo[0], o.x), d.x), tux = XDIV(XSUB(m.hi[0], o.x), d.x);
    float tly = XDIV(XSUB(m.lo[1], o.y), d.y), tuy = XDIV(XSUB(m.hi[1], o.y), d.y);
    float tlz = XDIV(XSUB(m.lo[2], o.z), d.z), tuz = XDIV(XSUB(m.hi[2], o.z), d.z);
    // Rust f32::min/max ignore NaN = fminf/fmaxf
    float t_min = fmaxf(fmaxf(fminf(tlx, tux), fminf(tly, tuy)), fminf(tlz, tuz));
    float t_max = fminf(fminf(fmaxf(tlx, tux), fmaxf(tly, tuy)), fmaxf(tlz, tuz));
    if (t_max < 0.0f) return false;
    if (t_min > t_max) return false;
    return true;
}

// ------------------------------------------------------------------ triangle.rs:189-241 (one lane of the AVX sweep)
// Returns true and t iff the reference's `has_intersect` mask is set for this triangle.
__device__ __forceinline__ bool tri_intersect(f3 v0, f3 e1, f3 e2, f3 o, f3 d, float& t_out) {
    f3 h = cross3(d, e2);
    float a = dot3(e1, h);
    if (-RBRT_MIN_DIST < a && a < RBRT_MIN_DIST) return false;          // c1 (two-sided, absolute)
    f3 s = o - v0;
    float f = XDIV(1.0f, a);                                             // true division (triangle.rs:203)
    float u = XMUL(f, dot3(s, h));
    if (u < 0.0f || u > 1.0f) return false;                              // c2
    f3 q = cross3(s, e1);
    float v = XMUL(f, dot3(d, q));
    if (v < 0.0f || XADD(u, v) > 1.0f) return false;                     // c3
    float t = XMUL(f, dot3(e2, q));
    if (!(t > RBRT_MIN_DIST && t < RBRT_T_CAP)) return false;            // c4
    t_out = t;
    return true;
}
// All compares are ordered-quiet like _CMP_LT_OQ/_CMP_GT_OQ (NaN -> false), so the early-outs above
// equal the reference's mask algebra has_intersect = !(c1|c2|c3) & c4 (triangle.rs:243-247) for every
// input, NaN included.

__device__ __forceinline__ void load_tri(const float4* __restrict__ tris, uint32_t i, f3& v0, f3& e1, f3& e2, uint32_t& orig) {
    float4 a = __ldg(tris + 3 * (size_t)i), b = __ldg(tris + 3 * (size_t)i + 1), c = __ldg(tris + 3 * (size_t)i + 2);
    v0 = mk3(a.x, a.y, a.z); e1 = mk3(b.x, b.y, b.z); e2 = mk3(c.x, c.y, c.z);
    orig = __float_as_uint(a.w);
}

// lexicographic (t, original index) minimum = "first index with the smallest t" (triangle.rs:392-410)
__device__ __forceinline__ void keep_min(float t, uint32_t orig, float& best_t, uint32_t& best_idx) {
    if (t < best_t || (t == best_t && orig < best_idx)) { best_t = t; best_idx = orig; }
}

// ------------------------------------------------------------------ brute force: the reference's own loop
__device__ __forceinline__ bool mesh_closest_brute(const SceneDev& S, const MeshDev& M, f3 o, f3 d,
                                                   float& best_t, uint32_t& best_idx, TraceCounters* cnt) {
    best_t = 1000000.0f; best_idx = 0xFFFFFFFFu;                         // min_param init (triangle.rs:398)
    for (uint32_t i = 0; i < M.n_tris; ++i) {
        f3 v0, e1, e2; uint32_t orig; float t;
        load_tri(S.tris, M.tri_base + i, v0, e1, e2, orig);
        if (tri_intersect(v0, e1, e2, o, d, t)) keep_min(t, orig, best_t, best_idx);
    }
    if (cnt) cnt->tris += M.n_tris;
    return best_idx != 0xFFFFFFFFu;
}

// ------------------------------------------------------------------ BVH2 traversal
// Node = 4 x float4 (64 B):
//   n0 = {c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y}
//   n1 = {c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y}
//   n2 = {c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z}
//   n3 = {bits(ref0), bits(ref1), -, -}     ref >= 0: node index; ref < 0: leaf (make_leaf_ref)
// Child boxes are padded at build time (bvh_build.cu) so that the FMA slab test below can never
// reject a box whose triangle the exact Moeller-Trumbore arithmetic above would accept.
__device__ __forceinline__ bool mesh_closest_bvh(const SceneDev& S, const MeshDev& M, f3 o, f3 d, float t_limit,
                                                 float& best_t, uint32_t& best_idx, TraceCounters* cnt) {
    best_t = 1000000.0f; best_idx = 0xFFFFFFFFu;
    // prune bound in t: nothing beyond min(best so far, caller's limit) can win; slack keeps the
    // prune conservative against the rounding of the exact test's t.
    float t_prune = t_limit;
    const float big = 1e30f;
    float idx = fabsf(d.x) > 1e-30f ? __fdividef(1.0f, d.x) : copysignf(big, d.x);
    float idy = fabsf(d.y) > 1e-30f ? __fdividef(1.0f, d.y) : copysignf(big, d.y);
    float idz = fabsf(d.z) > 1e-30f ? __fdividef(1.0f, d.z) : copysignf(big, d.z);
    float oox = o.x * idx, ooy = o.y * idy, ooz = o.z * idz;

    const float4* __restrict__ nodes = S.nodes + 4 * (size_t)M.node_base;
    int32_t stack[RBRT_STACK];
    int sp = 0;
    int32_t cur = M.root_ref;
    const int32_t SENTINEL = 0x7FFFFFFF;
    stack[sp++] = SENTINEL;
    uint32_t n_nodes = 0, n_tris = 0;

    while (cur != SENTINEL) {
        if (cur >= 0) {
            const float4* n = nodes + 4 * (size_t)cur;
            float4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
            ++n_nodes;
            float c0lox = __fmaf_rn(n0.x, idx, -oox), c0hix = __fmaf_rn(n0.y, idx, -oox);
            float c0loy = __fmaf_rn(n0.z, idy, -ooy), c0hiy = __fmaf_rn(n0.w, idy, -ooy);
            float c0loz = __fmaf_rn(n2.x, idz, -ooz), c0hiz = __fmaf_rn(n2.y, idz, -ooz);
            float c1lox = __fmaf_rn(n1.x, idx, -oox), c1hix = __fmaf_rn(n1.y, idx, -oox);
            float c1loy = __fmaf_rn(n1.z, idy, -ooy), c1hiy = __fmaf_rn(n1.w, idy, -ooy);
            float c1loz = __fmaf_rn(n2.z, idz, -ooz), c1hiz = __fmaf_rn(n2.w, idz, -ooz);
            float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
            float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), t_prune));
            float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
            float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), t_prune));
            bool h0 = t0n <= t0f, h1 = t1n <= t1f;
            int32_t r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if (h0 && h1) {
                bool swap = t1n < t0n;
                int32_t nearr = swap ? r1 : r0, farr = swap ? r0 : r1;
                stack[sp++] = farr;
                cur = nearr;
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else cur = stack[--sp];
        } else {
            uint32_t code = (uint32_t)(~cur);
            uint32_t first = code >> 3, count = (code & 7) + 1;
            for (uint32_t k = 0; k < count; ++k) {
                f3 v0, e1, e2; uint32_t orig; float t;
                load_tri(S.tris, M.tri_base + first + k, v0, e1, e2, orig);
                if (tri_intersect(v0, e1, e2, o, d, t)) {
                    keep_min(t, orig, best_t, best_idx);
                    t_prune = fminf(t_limit, __fmaf_rn(best_t, 1.0001f, 1e-4f));
                }
            }
            n_tris += count;
            cur = stack[--sp];
        }
    }
    if (cnt) { cnt->nodes += n_nodes; cnt->tris += n_tris; }
    return best_idx != 0xFFFFFFFFu;
}

// ------------------------------------------------------------------ scene.rs:19-43
template <bool BRUTE>
__device__ __forceinline__ Hit scene_hit(const SceneDev& S, f3 o, f3 d, TraceCounters* cnt) {
    Hit best; best.kind = -1; best.elem = 0; best.tri = 0; best.t = 0.0f; best.dist = 0.0f;
    float closest = 3.40282347e+38f;                                     // f32::MAX (scene.rs:21)
    for (uint32_t i = 0; i < S.n_spheres; ++i) {                         // spheres first, in order (scene.rs:23-31)
        float t, dist;
        int r = sphere_intersect(__ldg(S.spheres + i), o, d, t, dist);
        if (r < 0) { best.kind = -2; return best; }
        if (r && dist < closest) { closest = dist; best.kind = 0; best.elem = i; best.t = t; best.dist = dist; }
    }
    for (uint32_t mi = 0; mi < S.n_meshes; ++mi) {                       // then meshes, in order (scene.rs:33-41)
        const MeshDev& M = S.meshes[mi];
        if (M.n_tris == 0) continue;                                     // nothing the sweep could accept
        if (!mesh_bbox_hit(M, o, d)) continue;                           // mesh.rs:233
        float t; uint32_t idx; bool ok;
        if (BRUTE) ok = mesh_closest_brute(S, M, o, d, t, idx, cnt);
        else {
            // A mesh hit only matters if its dist beats `closest` (strict <).  dist is monotone in t
            // and ~ t*|d|; convert with a generous margin so the bound never cuts a winning hit.
            float t_limit = RBRT_T_CAP;
            if (closest < 3.0e38f) {
                float dl = len3(d);
                float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
                float lim = (closest * 1.001f + 1e-5f * (omax + closest) + 1e-6f) / dl;
                if (lim == lim) t_limit = fminf(t_limit, lim);
            }
            ok = mesh_closest_bvh(S, M, o, d, t_limit, t, idx, cnt);
        }
        if (!ok) continue;
        f3 p = o + t * d;                                                // ray.point_at (mesh.rs:247)
        float dist = len3(o - p);                                        // mesh.rs:248
        if (dist > RBRT_MIN_DIST && dist < RBRT_MAX_DIST && dist < closest) {   // mesh.rs:249 + scene.rs:36
            closest = dist; best.kind = 1; best.elem = mi; best.tri = idx; best.t = t; best.dist = dist;
        }
    }
    return best;
}

}  // namespace rbrt
