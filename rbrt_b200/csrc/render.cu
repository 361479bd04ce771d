// render.cu — the wavefront path tracer: render_scene + colorize (rbrt_lib/src/lib.rs:43-124).
//
// One batch = every pixel of this rank's shard x a run of consecutive samples.  Path id
// pid = s_local * P + j (j = pixel enumeration of the shard, 8x4 tiles, one warp = one tile).
// Per bounce iteration `it` (= depth 50-it of the reference's recursion):
//   k_trace : persistent warps pull 32 rays at a time from queue it&1, run Scene::hit, finish
//             missed paths (sky + attenuation product, written to out[pid]) and append hit slots to
//             one of three per-material index queues with ballot + one atomic per warp.
//   k_shade : persistent warps pull 32 hits of ONE material, run its scatter() and append the
//             continuation ray to queue (it+1)&1.
// After the last iteration k_accumulate adds the batch's per-path radiances to the per-pixel sums in
// sample order — the reference's `color += colorize(..)` (lib.rs:96-100) — with no float atomics,
// so an image is bit-reproducible and independent of scheduling.
#include <cstdio>
#include <cstdlib>
#include "engine.cuh"
#include "intersect.cuh"
#include "shade.cuh"

namespace rbrt {

// ------------------------------------------------------------------ warp helpers
// Every lane of a converged warp calls this; lanes with pred get consecutive slots.
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0) return 0;
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

__device__ __forceinline__ uint32_t warp_grab(uint32_t* head) {
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(head, 32u);
    return __shfl_sync(0xFFFFFFFFu, base, 0);
}

// ------------------------------------------------------------------ generate (cam.rs:64-82)
__global__ void __launch_bounds__(256) k_generate(WaveParams P) {
    uint32_t n_paths = P.s_count * P.paths_px;
    uint32_t stride = gridDim.x * blockDim.x;
    // n_paths is a multiple of 32 (paths_px is), so whole warps stay converged
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < n_paths; pid += stride) {
        uint32_t s_local = pid / P.paths_px, j = pid - s_local * P.paths_px;
        uint32_t row, col;
        bool valid = shard_pixel(P.sh, P.cam, j, row, col);
        f3 o = mk3(0, 0, 0), d = mk3(0, 0, 0);
        if (valid) {
            RngKey key; key.k0 = P.key0; key.k1 = P.key1;
            camera_ray(P.cam, row, col, key, row * P.cam.width + col, P.s_base + s_local, o, d);
        }
        uint32_t slot = warp_append(&P.ctr[0].ray_count, valid);
        if (valid) {
            P.q_o[0][slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pid));
            P.q_d[0][slot] = make_float4(d.x, d.y, d.z, 0.0f);
        }
    }
}

// ------------------------------------------------------------------ trace (scene.rs:19-43 + the miss arm of lib.rs:68-71)
template <bool BRUTE, bool COUNT>
__global__ void __launch_bounds__(256) k_trace(WaveParams P, uint32_t it) {
    IterCtr* c = P.ctr + it;
    const uint32_t n = c->ray_count;
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[ST_RAYS], (unsigned long long)n);
    const float4* __restrict__ qo = P.q_o[it & 1];
    const float4* __restrict__ qd = P.q_d[it & 1];
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    uint32_t nan_count = 0;
    for (;;) {
        uint32_t base = warp_grab(&c->ray_head);
        if (base >= n) break;
        uint32_t i = base + (threadIdx.x & 31);
        bool active = i < n;
        int mat_kind = -1;
        if (active) {
            float4 a = qo[i], b = qd[i];
            f3 o = mk3(a.x, a.y, a.z), d = mk3(b.x, b.y, b.z);
            uint32_t pid = __float_as_uint(a.w);
            Hit h = scene_hit<BRUTE>(P.S, o, d, COUNT ? &cnt : nullptr);
            if (h.kind >= 0) {
                if (it < P.max_depth) {                                   // depth > 0: scatter() will run (lib.rs:54)
                    uint32_t elem = h.kind == 0 ? h.elem : P.S.n_spheres + h.elem;
                    P.hit[i] = make_uint4(__float_as_uint(h.t), elem, h.tri, (uint32_t)h.kind);
                    mat_kind = (int)__ldg(P.S.mat_kind + elem);
                }                                                         // else: depth exhausted -> black (lib.rs:63-66)
            } else if (h.kind == -1) {
                // miss: sky, then unwind att_1 * (att_2 * (... * sky)) innermost first (lib.rs:62)
                f3 col = sky(d);
                for (int k = (int)it - 1; k >= 0; --k) {
                    uint32_t e = P.hist[(size_t)k * P.cap + pid];
                    float4 m = __ldg(P.S.mat + e);
                    f3 att = __ldg(P.S.mat_kind + e) == 2u ? mk3(1.0f, 1.0f, 1.0f) : mk3(m.x, m.y, m.z);
                    col = att * col;
                }
                P.out[pid] = make_float4(col.x, col.y, col.z, 0.0f);
            } else {
                ++nan_count;                                              // reference panics here (sphere.rs:33)
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t slot = warp_append(&c->mat_count[k], mat_kind == k);
            if (mat_kind == k) P.matq[k][slot] = i;
        }
    }
    if (nan_count) atomicAdd(&P.stats[ST_NAN], (unsigned long long)nan_count);
    if (COUNT) {
        atomicAdd(&P.stats[ST_NODES], (unsigned long long)cnt.nodes);
        atomicAdd(&P.stats[ST_TRIS], (unsigned long long)cnt.tris);
    }
}

// ------------------------------------------------------------------ shade (lib.rs:54-62 + the scatter impls)
__global__ void __launch_bounds__(256) k_shade(WaveParams P, uint32_t it) {
    IterCtr* c = P.ctr + it;
    const uint32_t n0 = c->mat_count[0], n1 = c->mat_count[1], n2 = c->mat_count[2];
    // virtual index space: each material's run is padded to a multiple of 32 so a warp never mixes kinds
    const uint32_t a0 = (n0 + 31u) & ~31u, a1 = a0 + ((n1 + 31u) & ~31u), total = a1 + ((n2 + 31u) & ~31u);
    if (total == 0) return;
    const float4* __restrict__ qo = P.q_o[it & 1];
    const float4* __restrict__ qd = P.q_d[it & 1];
    float4* __restrict__ no = P.q_o[(it + 1) & 1];
    float4* __restrict__ nd = P.q_d[(it + 1) & 1];
    RngKey key; key.k0 = P.key0; key.k1 = P.key1;
    for (;;) {
        uint32_t base = warp_grab(&c->shade_head);
        if (base >= total) break;
        uint32_t w = base + (threadIdx.x & 31);
        uint32_t kind, j, nk;
        if (w < a0) { kind = 0; j = w; nk = n0; }
        else if (w < a1) { kind = 1; j = w - a0; nk = n1; }
        else { kind = 2; j = w - a1; nk = n2; }
        bool active = j < nk;
        bool cont = false;
        f3 point = mk3(0, 0, 0), out_d = mk3(0, 0, 0);
        uint32_t pid = 0;
        if (active) {
            uint32_t i = P.matq[kind][j];
            float4 a = qo[i], b = qd[i];
            uint4 h = P.hit[i];
            f3 o = mk3(a.x, a.y, a.z), d = mk3(b.x, b.y, b.z);
            pid = __float_as_uint(a.w);
            float t = __uint_as_float(h.x);
            uint32_t elem = h.y;
            point = o + t * d;                                            // ray.point_at(t) (sphere.rs:49, mesh.rs:247)
            f3 normal;
            if (h.w == 0u) {                                              // sphere: p - c, un-normalised (sphere.rs:56)
                float4 s = __ldg(P.S.spheres + elem);
                normal = point - mk3(s.x, s.y, s.z);
            } else {                                                      // mesh: stored unit normal (mesh.rs:253-257)
                const MeshDev& M = P.S.meshes[elem - P.S.n_spheres];
                float4 nn = __ldg(P.S.normals + M.nrm_base + h.z);
                normal = mk3(nn.x, nn.y, nn.z);
            }
            uint32_t s_local = pid / P.paths_px, jp = pid - s_local * P.paths_px;
            uint32_t row, col;
            shard_pixel(P.sh, P.cam, jp, row, col);
            cont = scatter(kind, __ldg(P.S.mat + elem), d, point, normal, key, row * P.cam.width + col,
                           P.s_base + s_local, it + 1, out_d);
            if (cont) P.hist[(size_t)it * P.cap + pid] = (uint16_t)elem;
        }
        uint32_t slot = warp_append(&P.ctr[it + 1].ray_count, cont);
        if (cont) {
            no[slot] = make_float4(point.x, point.y, point.z, __uint_as_float(pid));
            nd[slot] = make_float4(out_d.x, out_d.y, out_d.z, 0.0f);
        }
    }
}

// ------------------------------------------------------------------ accumulate (lib.rs:95-100)
__global__ void __launch_bounds__(256) k_accumulate(WaveParams P, float4* __restrict__ accum) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.paths_px) return;
    uint32_t row, col;
    if (!shard_pixel(P.sh, P.cam, j, row, col)) return;
    size_t px = (size_t)row * P.cam.width + col;
    float4 acc = accum[px];
    for (uint32_t s = 0; s < P.s_count; ++s) {
        float4 c = P.out[(size_t)s * P.paths_px + j];
        acc.x = XADD(acc.x, c.x); acc.y = XADD(acc.y, c.y); acc.z = XADD(acc.z, c.z);
    }
    accum[px] = acc;
}

// ------------------------------------------------------------------ finalize (lib.rs:101,116-122)
__global__ void __launch_bounds__(256) k_finalize(const float4* __restrict__ accum, size_t n_px, float inv_spp,
                                                  uint8_t* __restrict__ rgb, float* __restrict__ hdr) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    float4 a = accum[i];
    f3 c = mk3(a.x, a.y, a.z) * inv_spp;
    if (hdr) { hdr[3 * i] = c.x; hdr[3 * i + 1] = c.y; hdr[3 * i + 2] = c.z; }
    if (rgb) {
        rgb[3 * i] = as_u8(XMUL(XSQRT(c.x), 256.0f));
        rgb[3 * i + 1] = as_u8(XMUL(XSQRT(c.y), 256.0f));
        rgb[3 * i + 2] = as_u8(XMUL(XSQRT(c.z), 256.0f));
    }
}

// ------------------------------------------------------------------ parity hooks
template <bool BRUTE>
__global__ void __launch_bounds__(256) k_trace_rays(SceneDev S, const rbrt_ray* __restrict__ rays, uint64_t n,
                                                    rbrt_hit* __restrict__ hits, unsigned long long* stats) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    if (i < n) {
        rbrt_ray r = rays[i];
        f3 o = mk3(r.origin.x, r.origin.y, r.origin.z), d = mk3(r.direction.x, r.direction.y, r.direction.z);
        Hit h = scene_hit<BRUTE>(S, o, d, stats ? &cnt : nullptr);
        rbrt_hit out;
        out.kind = h.kind >= 0 ? h.kind : RBRT_HIT_NONE;
        out.elem_idx = 0; out.tri_idx = 0; out.t = 0.0f; out.dist = 0.0f;
        out.point.x = out.point.y = out.point.z = 0.0f;
        out.normal.x = out.normal.y = out.normal.z = 0.0f;
        if (h.kind >= 0) {
            f3 p = o + h.t * d;
            f3 nrm;
            if (h.kind == 0) { float4 s = __ldg(S.spheres + h.elem); nrm = p - mk3(s.x, s.y, s.z); }
            else { float4 nn = __ldg(S.normals + S.meshes[h.elem].nrm_base + h.tri); nrm = mk3(nn.x, nn.y, nn.z); }
            out.elem_idx = h.elem; out.tri_idx = h.tri; out.t = h.t; out.dist = h.dist;
            out.point.x = p.x; out.point.y = p.y; out.point.z = p.z;
            out.normal.x = nrm.x; out.normal.y = nrm.y; out.normal.z = nrm.z;
        } else if (h.kind == -2 && stats) atomicAdd(&stats[ST_NAN], 1ull);
        hits[i] = out;
    }
    if (stats) {
        if (cnt.nodes) atomicAdd(&stats[ST_NODES], (unsigned long long)cnt.nodes);
        if (cnt.tris) atomicAdd(&stats[ST_TRIS], (unsigned long long)cnt.tris);
    }
}

__global__ void __launch_bounds__(256) k_primary_rays(CamDev cam, uint32_t key0, uint32_t key1, uint32_t sample, rbrt_ray* __restrict__ rays) {
    uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= cam.width * cam.height) return;
    uint32_t row = px / cam.width, col = px - row * cam.width;
    RngKey key; key.k0 = key0; key.k1 = key1;
    f3 o, d;
    camera_ray(cam, row, col, key, px, sample, o, d);
    rbrt_ray r; r.origin.x = o.x; r.origin.y = o.y; r.origin.z = o.z; r.direction.x = d.x; r.direction.y = d.y; r.direction.z = d.z;
    rays[px] = r;
}

// ====================================================================== host side
static CamDev make_cam(const rbrt_camera& c) {
    CamDev d;
    d.pos[0] = c.position.x; d.pos[1] = c.position.y; d.pos[2] = c.position.z;
    d.right[0] = c.right.x; d.right[1] = c.right.y; d.right[2] = c.right.z;
    d.up[0] = c.up.x; d.up[1] = c.up.y; d.up[2] = c.up.z;
    d.center[0] = c.img_center_point.x; d.center[1] = c.img_center_point.y; d.center[2] = c.img_center_point.z;
    d.mm_per_pix_hor = c.mm_per_pix_hor; d.mm_per_pix_vert = c.mm_per_pix_vert;
    d.width = c.img_width_pix; d.height = c.img_height_pix;
    return d;
}

void free_wave_buffers(WaveBuffers& wb) {
    for (int i = 0; i < 2; ++i) { cudaFree(wb.q_o[i]); cudaFree(wb.q_d[i]); }
    cudaFree(wb.hit); for (int i = 0; i < 3; ++i) cudaFree(wb.matq[i]);
    cudaFree(wb.out); cudaFree(wb.hist); cudaFree(wb.ctr); cudaFree(wb.stats); cudaFree(wb.accum);
    cudaFree(wb.rgb); cudaFree(wb.hdr);
    for (cudaEvent_t e : wb.ev) cudaEventDestroy(e);
    wb = WaveBuffers();
}

#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, #x); } while (0)

static int ensure_wave_buffers(WaveBuffers& wb, uint32_t cap, uint32_t depth) {
    if (wb.cap >= cap && wb.depth_cap >= depth && wb.stats) return RBRT_OK;
    float4* accum = wb.accum; size_t accum_px = wb.accum_px; uint8_t* rgb = wb.rgb; float* hdr = wb.hdr; size_t out_px = wb.out_px;
    std::vector<cudaEvent_t> ev; ev.swap(wb.ev);
    wb.accum = nullptr; wb.rgb = nullptr; wb.hdr = nullptr;
    free_wave_buffers(wb);
    wb.accum = accum; wb.accum_px = accum_px; wb.rgb = rgb; wb.hdr = hdr; wb.out_px = out_px; wb.ev.swap(ev);
    size_t b = 0;
    for (int i = 0; i < 2; ++i) { CKR(cudaMalloc(&wb.q_o[i], 16ull * cap)); CKR(cudaMalloc(&wb.q_d[i], 16ull * cap)); b += 32ull * cap; }
    CKR(cudaMalloc(&wb.hit, 16ull * cap)); b += 16ull * cap;
    for (int i = 0; i < 3; ++i) { CKR(cudaMalloc(&wb.matq[i], 4ull * cap)); b += 4ull * cap; }
    CKR(cudaMalloc(&wb.out, 16ull * cap)); b += 16ull * cap;
    CKR(cudaMalloc(&wb.hist, 2ull * cap * depth)); b += 2ull * cap * depth;
    CKR(cudaMalloc(&wb.ctr, sizeof(IterCtr) * (depth + 2)));
    CKR(cudaMalloc(&wb.stats, 8 * ST_COUNT));
    wb.cap = cap; wb.depth_cap = depth; wb.bytes = b;
    return RBRT_OK;
}

int render_accum(const Scene& sc, const rbrt_camera& cam, uint32_t spp, const rbrt_render_opts* opts,
                 float4* d_accum, cudaStream_t st, rbrt_stats* stats) {
    const uint32_t W = cam.img_width_pix, H = cam.img_height_pix;
    if (!W || !H || !spp) { set_error("empty image or zero samples"); return RBRT_E_INVALID; }
    if ((uint64_t)W * H > 0x7FFFFFFFull) { set_error("image too large"); return RBRT_E_INVALID; }
    rbrt_render_opts o{}; if (opts) o = *opts;
    const uint32_t max_depth = o.max_depth ? o.max_depth : 50;
    if (max_depth > 1024) { set_error("max_depth > 1024"); return RBRT_E_INVALID; }
    if (o.integrator != 0) { set_error("unknown integrator %u", o.integrator); return RBRT_E_INVALID; }
    ShardDev sh;
    sh.rank = 0; sh.count = 1; sh.s0 = 0; sh.s1 = spp;
    sh.tiles_x = (W + 7) / 8;
    sh.tiles_total = sh.tiles_x * ((H + 3) / 4);
    if (o.shard_count > 1) {
        if (o.shard_rank >= o.shard_count) { set_error("shard_rank >= shard_count"); return RBRT_E_INVALID; }
        if (o.shard_mode == RBRT_SHARD_TILES) { sh.rank = o.shard_rank; sh.count = o.shard_count; }
        else if (o.shard_mode == RBRT_SHARD_SAMPLES) {
            sh.s0 = (uint32_t)((uint64_t)spp * o.shard_rank / o.shard_count);
            sh.s1 = (uint32_t)((uint64_t)spp * (o.shard_rank + 1) / o.shard_count);
        } else { set_error("shard_count > 1 needs shard_mode TILES or SAMPLES"); return RBRT_E_INVALID; }
    }
    sh.tiles_mine = sh.tiles_total > sh.rank ? (sh.tiles_total - sh.rank + sh.count - 1) / sh.count : 0;
    const uint32_t P = sh.tiles_mine * 32;
    cudaEvent_t ev0, ev1;
    CKR(cudaEventCreate(&ev0)); CKR(cudaEventCreate(&ev1));
    CKR(cudaMemsetAsync(d_accum, 0, 16ull * W * H, st));
    uint32_t launches = 0, iterations = 0;
    WaveBuffers& wb = sc.wb;
    CKR(cudaEventRecord(ev0, st));
    if (P && sh.s1 > sh.s0) {
        // Paths in flight per batch.  Every bounce iteration is one trace + one shade launch whose duration is
        // bounded below by its slowest ray, so the ~45 sparsely populated tail iterations cost the same for a
        // small batch as for a large one: the default is therefore "as many paths as fit" — up to 2^27 paths
        // (108 + 2*max_depth bytes of wavefront state each: 27.9 GB at depth 50) and at most half of the free HBM.
        uint32_t target = o.batch_paths;
        if (!target) {
            size_t free_b = 0, total_b = 0;
            CKR(cudaMemGetInfo(&free_b, &total_b));
            free_b += wb.bytes;                                           // what a re-allocation would release first
            uint64_t per_path = 108ull + 2ull * max_depth;
            uint64_t fit = (free_b / 2) / per_path;
            target = (uint32_t)(fit < (1ull << 21) ? (1ull << 21) : (fit > (1ull << 27) ? (1ull << 27) : fit));
        }
        uint32_t S_b = target / P; if (S_b < 1) S_b = 1; if (S_b > sh.s1 - sh.s0) S_b = sh.s1 - sh.s0;
        if ((uint64_t)S_b * P > 0x7FFFFFFFull) { set_error("batch too large"); return RBRT_E_INVALID; }
        const uint32_t cap = S_b * P;
        int rc = ensure_wave_buffers(wb, cap, max_depth);
        if (rc) return rc;
        CKR(cudaMemsetAsync(wb.stats, 0, 8 * ST_COUNT, st));
        WaveParams wp;
        wp.S = sc.dev; wp.cam = make_cam(cam); wp.sh = sh;
        wp.key0 = (uint32_t)o.seed; wp.key1 = (uint32_t)(o.seed >> 32);
        wp.cap = wb.cap; wp.paths_px = P; wp.max_depth = max_depth;
        for (int i = 0; i < 2; ++i) { wp.q_o[i] = wb.q_o[i]; wp.q_d[i] = wb.q_d[i]; }
        wp.hit = wb.hit; for (int i = 0; i < 3; ++i) wp.matq[i] = wb.matq[i];
        wp.out = wb.out; wp.hist = wb.hist; wp.ctr = wb.ctr; wp.stats = wb.stats;
        const int grid = sc.sm_count * 8;
        const bool brute = o.trace_mode == RBRT_TRACE_BRUTE;
        const bool count = (o.flags & RBRT_OPT_COUNT_VISITS) != 0;
        const bool time_kernels = (o.flags & RBRT_OPT_TIME_KERNELS) != 0;      // bracket every trace launch with events -> stats.ms_trace
        size_t ev_used = 0;
        auto next_event = [&]() -> cudaEvent_t {
            if (ev_used == wb.ev.size()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); wb.ev.push_back(e); }
            return wb.ev[ev_used++];
        };
        for (uint32_t s_base = sh.s0; s_base < sh.s1; s_base += S_b) {
            wp.s_base = s_base; wp.s_count = (sh.s1 - s_base < S_b) ? sh.s1 - s_base : S_b;
            CKR(cudaMemsetAsync(wb.ctr, 0, sizeof(IterCtr) * (max_depth + 2), st));
            CKR(cudaMemsetAsync(wb.out, 0, 16ull * wp.s_count * P, st));
            k_generate<<<grid, 256, 0, st>>>(wp); ++launches;
            for (uint32_t it = 0; it <= max_depth; ++it) {
                if (time_kernels) CKR(cudaEventRecord(next_event(), st));
                if (brute) { if (count) k_trace<true, true><<<grid, 256, 0, st>>>(wp, it); else k_trace<true, false><<<grid, 256, 0, st>>>(wp, it); }
                else { if (count) k_trace<false, true><<<grid, 256, 0, st>>>(wp, it); else k_trace<false, false><<<grid, 256, 0, st>>>(wp, it); }
                ++launches; ++iterations;
                if (time_kernels) CKR(cudaEventRecord(next_event(), st));
                if (it < max_depth) { k_shade<<<grid, 256, 0, st>>>(wp, it); ++launches; }
            }
            k_accumulate<<<(P + 255) / 256, 256, 0, st>>>(wp, d_accum); ++launches;
            CKR(cudaGetLastError());
        }
    }
    CKR(cudaEventRecord(ev1, st));
    if (stats) {
        CKR(cudaEventSynchronize(ev1));
        float ms = 0; CKR(cudaEventElapsedTime(&ms, ev0, ev1));
        unsigned long long h[ST_COUNT] = {0, 0, 0, 0};
        if (wb.stats && P && sh.s1 > sh.s0) CKR(cudaMemcpy(h, wb.stats, sizeof(h), cudaMemcpyDeviceToHost));
        stats->rays = h[ST_RAYS]; stats->nan_rays = h[ST_NAN]; stats->node_visits = h[ST_NODES]; stats->tri_tests = h[ST_TRIS];
        uint64_t valid_px = 0;
        for (uint32_t tj = 0; tj < sh.tiles_mine; ++tj) {
            uint32_t T = tj * sh.count + sh.rank, ty = T / sh.tiles_x, tx = T - ty * sh.tiles_x;
            uint32_t w = W - tx * 8 < 8 ? W - tx * 8 : 8, hh = H - ty * 4 < 4 ? H - ty * 4 : 4;
            valid_px += (uint64_t)w * hh;
        }
        stats->paths = valid_px * (sh.s1 - sh.s0);
        stats->ms_device = ms; stats->launches = launches; stats->iterations = iterations;
        if (P && sh.s1 > sh.s0 && (o.flags & RBRT_OPT_TIME_KERNELS)) {
            double tr = 0;
            for (size_t i = 0; i + 1 < wb.ev.size() && i + 1 < 2ull * iterations; i += 2) {
                float t = 0; CKR(cudaEventElapsedTime(&t, wb.ev[i], wb.ev[i + 1])); tr += t;
            }
            stats->ms_trace = tr;
            if (getenv("RBRT_DEBUG_ITERS")) {                              // per-iteration trace time + queue sizes of the LAST batch
                std::vector<IterCtr> hc(max_depth + 2);
                CKR(cudaMemcpy(hc.data(), wb.ctr, sizeof(IterCtr) * (max_depth + 2), cudaMemcpyDeviceToHost));
                size_t per_batch = max_depth + 1, first = 2 * (iterations - per_batch);
                for (uint32_t it = 0; it <= max_depth; ++it) {
                    float t = 0; cudaEventElapsedTime(&t, wb.ev[first + 2 * it], wb.ev[first + 2 * it + 1]);
                    fprintf(stderr, "it %2u rays %9u  lambert %9u metal %9u glass %9u  trace %8.3f ms\n", it, hc[it].ray_count,
                            hc[it].mat_count[0], hc[it].mat_count[1], hc[it].mat_count[2], t);
                }
            }
        }
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    return RBRT_OK;
}

int finalize(const float4* d_accum, uint32_t W, uint32_t H, uint32_t spp, uint8_t* d_rgb, float* d_hdr, cudaStream_t st) {
    size_t n = (size_t)W * H;
    if (!n || !spp) { set_error("empty image or zero samples"); return RBRT_E_INVALID; }
    float inv = 1.0f / (float)spp;                                        // lib.rs:101
    k_finalize<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_accum, n, inv, d_rgb, d_hdr);
    CKR(cudaGetLastError());
    return RBRT_OK;
}

int trace_rays_device(const Scene& sc, const rbrt_ray* d_rays, uint64_t n, uint32_t mode, rbrt_hit* d_hits,
                      unsigned long long* d_stats, cudaStream_t st) {
    if (!n) return RBRT_OK;
    unsigned g = (unsigned)((n + 255) / 256);
    if (mode == RBRT_TRACE_BRUTE) k_trace_rays<true><<<g, 256, 0, st>>>(sc.dev, d_rays, n, d_hits, d_stats);
    else k_trace_rays<false><<<g, 256, 0, st>>>(sc.dev, d_rays, n, d_hits, d_stats);
    CKR(cudaGetLastError());
    return RBRT_OK;
}

int primary_rays_device(const rbrt_camera& cam, uint64_t seed, uint32_t sample, rbrt_ray* d_rays, cudaStream_t st) {
    size_t n = (size_t)cam.img_width_pix * cam.img_height_pix;
    if (!n) return RBRT_OK;
    k_primary_rays<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(make_cam(cam), (uint32_t)seed, (uint32_t)(seed >> 32), sample, d_rays);
    CKR(cudaGetLastError());
    return RBRT_OK;
}

}  // namespace rbrt
