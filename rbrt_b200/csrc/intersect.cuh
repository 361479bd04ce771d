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

// ------------------------------------------------------------------ triangle.rs:92-130 + 412-441 (BasicTriangle element)
// returns 1 hit, 0 miss.  Same op sequence as the mesh sweep, but: `u` must be CONTAINED in [0,1] (a NaN u rejects,
// `!(0.0..=1.0).contains(&u)`), no upper cap on t, dist accepted in [min_dist, max_dist] inclusive (NaN dist passes).
__device__ __forceinline__ int basic_triangle_intersect(const float4* __restrict__ rec, f3 o, f3 d, float& t_out, float& dist_out) {
    float4 a4 = __ldg(rec), b4 = __ldg(rec + 1), c4 = __ldg(rec + 2);
    f3 v0 = mk3(a4.x, a4.y, a4.z), e1 = mk3(b4.x, b4.y, b4.z), e2 = mk3(c4.x, c4.y, c4.z);
    f3 h = cross3(d, e2);
    float a = dot3(e1, h);
    if (-RBRT_MIN_DIST < a && a < RBRT_MIN_DIST) return 0;
    float f = XDIV(1.0f, a);
    f3 s = o - v0;
    float u = XMUL(f, dot3(s, h));
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    f3 q = cross3(s, e1);
    float v = XMUL(f, dot3(d, q));
    if (v < 0.0f || XADD(u, v) > 1.0f) return 0;
    float t = XMUL(f, dot3(e2, q));
    if (!(t > RBRT_MIN_DIST)) return 0;
    f3 p = o + t * d;
    float dist = len3(o - p);
    if (dist < RBRT_MIN_DIST || dist > RBRT_MAX_DIST) return 0;
    t_out = t; dist_out = dist;
    return 1;
}

// One entry of Scene.elements: sphere or BasicTriangle.  returns 1 hit, 0 miss, -1 NaN discriminant (spheres only)
__device__ __forceinline__ int element_intersect(const SceneDev& S, uint32_t i, f3 o, f3 d, float& t_out, float& dist_out) {
    float4 e = __ldg(S.spheres + i);
    if (S.n_etris && __ldg(S.elem_kind + i)) return basic_triangle_intersect(S.etris + 4 * (size_t)__float_as_uint(e.x), o, d, t_out, dist_out);
    return sphere_intersect(e, o, d, t_out, dist_out);
}

// hit_normal of an element hit: sphere p - c, un-normalised (sphere.rs:56); BasicTriangle its stored unit normal (triangle.rs:433)
__device__ __forceinline__ f3 element_normal(const SceneDev& S, uint32_t i, f3 p) {
    float4 e = __ldg(S.spheres + i);
    if (S.n_etris && __ldg(S.elem_kind + i)) { float4 n = __ldg(S.etris + 4 * (size_t)__float_as_uint(e.x) + 3); return mk3(n.x, n.y, n.z); }
    return p - mk3(e.x, e.y, e.z);
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

// The same test without early-outs — literally the reference's lane: every quantity is computed, then the four masks
// are combined (triangle.rs:189-247).  Used by the warp-voted traversal, where the lanes of a step run in lock-step
// anyway: no divergence, and the three record loads are issued together instead of v0 waiting behind the first cull.
__device__ __forceinline__ bool tri_intersect_masks(f3 v0, f3 e1, f3 e2, f3 o, f3 d, float& t_out) {
    f3 h = cross3(d, e2);
    float a = dot3(e1, h);
    f3 s = o - v0;
    float f = XDIV(1.0f, a);                                             // a == 0 -> inf, masked by c1 like the AVX lane
    float u = XMUL(f, dot3(s, h));
    f3 q = cross3(s, e1);
    float v = XMUL(f, dot3(d, q));
    float t = XMUL(f, dot3(e2, q));
    const bool c1 = (-RBRT_MIN_DIST < a) & (a < RBRT_MIN_DIST);
    const bool c2 = (u < 0.0f) | (u > 1.0f);
    const bool c3 = (v < 0.0f) | (XADD(u, v) > 1.0f);
    const bool c4 = (t > RBRT_MIN_DIST) & (t < RBRT_T_CAP);
    t_out = t;
    return !(c1 | c2 | c3) & c4;
}

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

// ------------------------------------------------------------------ 4-wide BVH traversal, 64-byte nodes
// Node = 4 x uint4: twelve words {child k: x, y, z}, each word = qlo | qhi << 16 on the mesh's 16-bit grid, then the four
// child refs (ref >= 0: node index; ref < 0: leaf, make_leaf_ref); unused slots hold an inverted box.
//   w0 = {c0.x, c0.y, c0.z, c1.x}  w1 = {c1.y, c1.z, c2.x, c2.y}  w2 = {c2.z, c3.x, c3.y, c3.z}  w3 = {ref0..ref3}
// The trace kernel is bound by the LATENCY of dependent node fetches (profiles/), so the tree is 4 wide: half as
// many dependent visits per ray as the binary tree it is collapsed from, for the same 16 bytes per child.
// A 16-bit value becomes the float 2^23 + q with ONE byte-permute (PRMT builds 0x4B00hhll); the permute selector
// also picks the near or the far bound for the ray's direction sign, so there is no per-axis min/max either:
//   t = fma(2^23 + q, A, B')  with  A = step * (1/d),  B' = (org - o) * (1/d) - 2^23 * A.
// B' is rounded once at magnitude ~2^23 |A|, i.e. by up to half a grid step of t; bvh_build.cu therefore quantises
// child boxes outward AND adds one more step of margin, so the slab test can never reject a box whose triangle
// the exact Moeller-Trumbore arithmetic above would accept.
#define RBRT_SENTINEL 0x7FFFFFFF
#define RBRT_SEL_LO 0x7610u           // PRMT selectors: bytes {0,1} / {2,3} of the node word under 0x4B00....
#define RBRT_SEL_HI 0x7632u
#define RBRT_MISS_T 3.0e38f

struct RaySlabs {
    float ax, ay, az, bx, by, bz;     // per-axis t = fma(m, a, b)
    uint32_t nx, ny, nz;              // selector of the NEAR bound per axis (far = near ^ 0x22)
    uint32_t fx, fy, fz;              // selectors of the FAR bounds, kept in registers too (the compiler otherwise re-derives all six
                                      // from three packed values at every node visit: 6 ALU-pipe instructions per visit, measured)
};

__device__ __forceinline__ RaySlabs ray_slabs(const MeshDev& M, f3 o, f3 d) {
    const float big = 1e25f;
    float idx = fabsf(d.x) > 1e-25f ? __fdividef(1.0f, d.x) : copysignf(big, d.x);
    float idy = fabsf(d.y) > 1e-25f ? __fdividef(1.0f, d.y) : copysignf(big, d.y);
    float idz = fabsf(d.z) > 1e-25f ? __fdividef(1.0f, d.z) : copysignf(big, d.z);
    RaySlabs r;
    r.ax = M.qstep[0] * idx; r.ay = M.qstep[1] * idy; r.az = M.qstep[2] * idz;
    r.bx = __fmaf_rn(-8388608.0f, r.ax, (M.qorg[0] - o.x) * idx);
    r.by = __fmaf_rn(-8388608.0f, r.ay, (M.qorg[1] - o.y) * idy);
    r.bz = __fmaf_rn(-8388608.0f, r.az, (M.qorg[2] - o.z) * idz);
    r.nx = idx >= 0.0f ? RBRT_SEL_LO : RBRT_SEL_HI;
    r.ny = idy >= 0.0f ? RBRT_SEL_LO : RBRT_SEL_HI;
    r.nz = idz >= 0.0f ? RBRT_SEL_LO : RBRT_SEL_HI;
    r.fx = r.nx ^ 0x22u; r.fy = r.ny ^ 0x22u; r.fz = r.nz ^ 0x22u;
    asm volatile("" : "+r"(r.nx), "+r"(r.ny), "+r"(r.nz), "+r"(r.fx), "+r"(r.fy), "+r"(r.fz));   // opaque: keep them materialised
    return r;
}

// (PTX prmt directly: __byte_perm masks its selector with 0x7777 first, and because the selectors are deliberately opaque to the
// compiler that mask was re-applied at every node visit — six LOP3 per visit on the pipe that binds the kernel)
__device__ __forceinline__ float q16(uint32_t w, uint32_t sel) {
    uint32_t r;
#ifdef __CUDA_ARCH__
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(sel));
#else                                                                     // host build of this header (tests/host_device): what PRMT does, byte by byte
    const uint64_t pool = ((uint64_t)0x4B000000u << 32) | w;
    r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
#endif
    return __uint_as_float(r);
}

// One node visit: returns the next reference to process (nearest hit child, or the popped stack top); the other
// hit children are pushed far-to-near.
// Traversal stacks.  StackL: plain per-lane array (local memory: entry e of all lanes shares a 128-byte line, so lanes at
// different depths touch different lines — up to 32 L1 requests per push).  StackS: the first K entries live in shared
// memory, laid out [entry][thread] (bank = lane, conflict-free at any mix of depths: one request per push), deeper
// entries spill to the local array.
struct StackL {
    int32_t* l;
    __device__ __forceinline__ void put(int i, int32_t v) const { l[i] = v; }
    __device__ __forceinline__ int32_t get(int i) const { return l[i]; }
};
template <int K, int STRIDE>
struct StackS {
    int32_t* s;      // &smem[0][threadIdx.x]
    int32_t* l;
    __device__ __forceinline__ void put(int i, int32_t v) const { if (i < K) s[i * STRIDE] = v; else l[i - K] = v; }
    __device__ __forceinline__ int32_t get(int i) const { return i < K ? s[i * STRIDE] : l[i - K]; }
};

// 256-bit read-only load (sm_100: LDG.E.256): a 64-byte node is two requests to L1 instead of four (the trace kernel runs
// L1TEX at 75-80 % of its peak, mostly on these gathers).  p must be 32-byte aligned.
// L2::evict_last: the nodes (C3: ~40 MB, every ray gathers from them) compete in L2 with the path records streaming through (4.3 GB per
// batch) and with the triangle records; k_trace's L2 hit rate is 62 % and its warps mostly wait on these loads.  Measured on C3: trace
// 22.65 -> 22.47 ms, frame 34.76 -> 34.51 ms (the same hint on the triangle records: 22.24 -> 22.41 ms, not adopted).  The other hints tried made things slower or changed nothing — streaming (.cs) record loads
// / stores, evict_last for triangles, L1::evict_last for nodes, L1::no_allocate for triangles or records (profiles/r2_exp_l2_hints.jsonl).
#ifndef RBRT_NODE_L2_EVICT_LAST
#define RBRT_NODE_L2_EVICT_LAST 1
#endif
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
#if RBRT_NODE_L2_EVICT_LAST
    asm volatile("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// PF (experiment, off: -DRBRT_TAIL_PREFETCH turns it on in the tail kernel): as soon as the child references are there, the lines of ALL
// internal children and of the first triangle of leaf children are prefetched, before the slab tests decide which one is visited next.
// Measured SLOWER (tail kernel 1.3 -> 2.3 ms on a 1/8 shard of C3), see render.cu.
template <bool PF, class STK>
__device__ __forceinline__ int32_t bvh4_step(const uint4* __restrict__ nd, const uint4* __restrict__ nodes, const float4* __restrict__ tris_m,
                                             const RaySlabs& R, float t_prune, const STK& stack, int& sp) {
#ifndef RBRT_LDG128
    uint4 w0, w1, w2, w3;
    ldg256(nd, w0, w1); ldg256(nd + 2, w2, w3);
#else
    const uint4 w0 = __ldg(nd), w1 = __ldg(nd + 1), w2 = __ldg(nd + 2), w3 = __ldg(nd + 3);
#endif
    if (PF) {
        const int32_t cr[4] = {(int32_t)w3.x, (int32_t)w3.y, (int32_t)w3.z, (int32_t)w3.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const void* a = cr[k] >= 0 ? (const void*)(nodes + 4 * (size_t)cr[k]) : (const void*)(tris_m + 3 * (size_t)(((uint32_t)~cr[k]) >> 3));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(a));
        }
    }
    const uint32_t fx = R.fx, fy = R.fy, fz = R.fz;
    float t[4]; int32_t r[4];
#define RBRT_CHILD(k, X, Y, Z) { \
        float tn = fmaxf(fmaxf(__fmaf_rn(q16(X, R.nx), R.ax, R.bx), __fmaf_rn(q16(Y, R.ny), R.ay, R.by)), fmaxf(__fmaf_rn(q16(Z, R.nz), R.az, R.bz), 0.0f)); \
        float tf = fminf(fminf(__fmaf_rn(q16(X, fx), R.ax, R.bx), __fmaf_rn(q16(Y, fy), R.ay, R.by)), fminf(__fmaf_rn(q16(Z, fz), R.az, R.bz), t_prune)); \
        t[k] = tn <= tf ? tn : RBRT_MISS_T; }
    RBRT_CHILD(0, w0.x, w0.y, w0.z) RBRT_CHILD(1, w0.w, w1.x, w1.y) RBRT_CHILD(2, w1.z, w1.w, w2.x) RBRT_CHILD(3, w2.y, w2.z, w2.w)
#undef RBRT_CHILD
    r[0] = (int32_t)w3.x; r[1] = (int32_t)w3.y; r[2] = (int32_t)w3.z; r[3] = (int32_t)w3.w;
#ifdef RBRT_FULL_SORT
    // sort the four (entry distance, ref) pairs ascending; misses end up last (measured 1-3 % slower than nearest-only on C2/C3/C4)
#define RBRT_CSWAP(a, b) { const bool sw = t[b] < t[a]; const float tl = fminf(t[a], t[b]), th = fmaxf(t[a], t[b]); \
        const int32_t ra = sw ? r[b] : r[a], rb = sw ? r[a] : r[b]; t[a] = tl; t[b] = th; r[a] = ra; r[b] = rb; }
    RBRT_CSWAP(0, 1) RBRT_CSWAP(2, 3) RBRT_CSWAP(0, 2) RBRT_CSWAP(1, 3) RBRT_CSWAP(1, 2)
#undef RBRT_CSWAP
#else
    // bring the nearest child to slot 0 (three exchanges); the other hit children are pushed in slot order
#define RBRT_CSWAP(a, b) { const bool sw = t[b] < t[a]; const float tl = fminf(t[a], t[b]), th = fmaxf(t[a], t[b]); \
        const int32_t ra = sw ? r[b] : r[a], rb = sw ? r[a] : r[b]; t[a] = tl; t[b] = th; r[a] = ra; r[b] = rb; }
    RBRT_CSWAP(0, 1) RBRT_CSWAP(2, 3) RBRT_CSWAP(0, 2)
#undef RBRT_CSWAP
#endif
    if (t[3] < RBRT_MISS_T) stack.put(sp++, r[3]);
    if (t[2] < RBRT_MISS_T) stack.put(sp++, r[2]);
    if (t[1] < RBRT_MISS_T) stack.put(sp++, r[1]);
    return t[0] < RBRT_MISS_T ? r[0] : stack.get(--sp);
}

// Leaf: <= 8 contiguous triangle records, exact test, running lexicographic minimum.
__device__ __forceinline__ void leaf_step(const float4* __restrict__ tris, uint32_t tri_base, int32_t cur, f3 o, f3 d, float t_limit,
                                          float& best_t, uint32_t& best_idx, float& t_prune, uint32_t& n_tris) {
    uint32_t code = (uint32_t)(~cur);
    uint32_t first = code >> 3, count = (code & 7) + 1;
    for (uint32_t k = 0; k < count; ++k) {
        f3 v0, e1, e2; uint32_t orig; float t;
        load_tri(tris, tri_base + first + k, v0, e1, e2, orig);
        if (tri_intersect(v0, e1, e2, o, d, t)) {
            keep_min(t, orig, best_t, best_idx);
            // prune bound in t: nothing beyond min(best so far, caller's limit) can win; the slack keeps the prune
            // conservative against the rounding of the exact test's t.
            t_prune = fminf(t_limit, __fmaf_rn(best_t, 1.0001f, 1e-4f));
        }
    }
    n_tris += count;
}

// ONE triangle of the current leaf (warp-voted traversal): test it, then `cur` becomes the rest of the leaf or the popped
// stack top.  Same arithmetic and the same order of triangles within a leaf as leaf_step.
template <class STK>
__device__ __forceinline__ void leaf_step_one(const float4* __restrict__ tris, uint32_t tri_base, int32_t& cur, f3 o, f3 d, float t_limit,
                                              float& best_t, uint32_t& best_idx, float& t_prune, const STK& stack, int& sp) {
    const uint32_t code = (uint32_t)(~cur);
    f3 v0, e1, e2; uint32_t orig; float t;
    load_tri(tris, tri_base + (code >> 3), v0, e1, e2, orig);
    if (tri_intersect_masks(v0, e1, e2, o, d, t)) {
        keep_min(t, orig, best_t, best_idx);
        t_prune = fminf(t_limit, __fmaf_rn(best_t, 1.0001f, 1e-4f));
    }
    cur = (code & 7u) ? (int32_t)~(code + 7u) : stack.get(--sp);              // {first + 1, count - 1}: +8 on first, -1 on the count field
}

// Warp-voted traversal (k_trace, k_tail).  The while-while form ("every lane descends to its next leaf, then the leaves
// are processed together") runs the node loop until the SLOWEST lane has found a leaf: measured 11 of 32 lanes per
// instruction there.  Here every step is ONE node visit or ONE triangle test, and the warp runs the kind that more of
// its lanes are waiting for; lanes of the other kind sit the step out, so a step always serves at least half of the
// lanes that have work.  A lane's own sequence of visits and tests is unchanged, hence so is every result bit.
// Leaves the loop when fewer than `threshold` lanes still have work (the caller re-fills lanes from the queue).
#ifndef RBRT_VOTE_N
#define RBRT_VOTE_N 1      // node step iff  lanes at nodes * RBRT_VOTE_N >= lanes at leaves * RBRT_VOTE_L
#define RBRT_VOTE_L 1
#endif
template <bool COUNT, bool PF, class STK>
__device__ __forceinline__ void traverse_voted(const uint4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t tri_base,
                                               const RaySlabs& R, f3 o, f3 d, float t_limit, const STK& stack, int& sp, int32_t& cur,
                                               float& best_t, uint32_t& best_idx, float& t_prune, int threshold,
                                               uint32_t& n_nodes, uint32_t& n_tris) {
    asm volatile("" : "+r"(threshold));                                   // keep it in a register (otherwise re-derived from three values every step)
    for (;;) {
        const bool at_node = (uint32_t)cur < (uint32_t)RBRT_SENTINEL, at_leaf = cur < 0;
        const int nn = __popc(__ballot_sync(0xFFFFFFFFu, at_node)), nl = __popc(__ballot_sync(0xFFFFFFFFu, at_leaf));
        if (nn + nl < threshold) break;
        if (nn * RBRT_VOTE_N >= nl * RBRT_VOTE_L) {
            if (at_node) { cur = bvh4_step<PF>(nodes + 4 * (size_t)cur, nodes, tris + 3 * (size_t)tri_base, R, t_prune, stack, sp); if (COUNT) ++n_nodes; }
        } else if (at_leaf) {
            leaf_step_one(tris, tri_base, cur, o, d, t_limit, best_t, best_idx, t_prune, stack, sp);
            if (COUNT) ++n_tris;
        }
    }
}

// Simple one-lane-one-ray traversal (parity hook, brute/finish kernels); the wavefront trace kernel (render.cu) runs the
// same steps warp-voted with dynamic fetch.
__device__ __forceinline__ bool mesh_closest_bvh(const SceneDev& S, const MeshDev& M, f3 o, f3 d, float t_limit,
                                                 float& best_t, uint32_t& best_idx, TraceCounters* cnt) {
    best_t = 1000000.0f; best_idx = 0xFFFFFFFFu;                         // min_param init (triangle.rs:398)
    float t_prune = t_limit;
    const RaySlabs R = ray_slabs(M, o, d);
    const uint4* __restrict__ nodes = reinterpret_cast<const uint4*>(S.nodes) + 4 * (size_t)M.node_base;
    int32_t lstack[RBRT_STACK];
    const StackL stack = {lstack};
    int sp = 0;
    int32_t cur = M.root_ref;
    stack.put(sp++, RBRT_SENTINEL);
    uint32_t n_nodes = 0, n_tris = 0;
    while (cur != RBRT_SENTINEL) {
        if (cur >= 0) { cur = bvh4_step<false>(nodes + 4 * (size_t)cur, nodes, S.tris, R, t_prune, stack, sp); ++n_nodes; }
        else { leaf_step(S.tris, M.tri_base, cur, o, d, t_limit, best_t, best_idx, t_prune, n_tris); cur = stack.get(--sp); }
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
        int r = element_intersect(S, i, o, d, t, dist);
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
