// host_math.cpp — rbrt_lib host arithmetic that feeds the hot path and must be bit-identical to the
// reference: Camera::new (cam.rs:22-62).  (The per-vertex mesh transform of load_mesh_vertices_from_file,
// mesh.rs:102-112 / vec3.rs:139-155, is in obj_loader.cpp.)  Plain f32, compiled with
// -ffp-contract=off so no a*b+c is fused (Rust never contracts).
#include <math.h>
#include "../../include/rbrt_gpu.h"

namespace {
struct V { float x, y, z; };
inline float len(V a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }               // vec3.rs:111-113
inline V nrm(V a) { float l = len(a); return V{a.x / l, a.y / l, a.z / l}; }              // vec3.rs:119-126
inline V crs(V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
}  // namespace

extern "C" int rbrt_camera_new(rbrt_vec3 position, rbrt_vec3 look_at, rbrt_vec3 up, uint32_t img_height_pix,
                               uint32_t img_width_pix, float focal_len_mm, rbrt_camera* out) {
    if (!out) return RBRT_E_INVALID;
    V la = nrm(V{look_at.x, look_at.y, look_at.z});
    V right = nrm(crs(la, nrm(V{up.x, up.y, up.z})));                                     // cam.rs:30-33
    float img_width_mm = 35.0f;                                                           // cam.rs:36
    float mm_per_pix_hor = img_width_mm / (float)img_width_pix;
    float img_height_mm = (float)img_height_pix * mm_per_pix_hor;
    float mm_per_pix_vert = img_height_mm / (float)img_height_pix;
    float k = focal_len_mm / 1000.0f;                                                     // cam.rs:42
    out->hor_fov_rad = 2.0f * atanf(2.0f * focal_len_mm / img_width_mm);
    out->img_width_pix = img_width_pix;
    out->img_height_mm = img_height_mm;
    out->vert_fov_rad = 2.0f * atanf(2.0f * focal_len_mm / img_height_mm);
    out->img_height_pix = img_height_pix;
    out->img_width_mm = img_width_mm;
    out->position = position;
    out->focal_len_mm = focal_len_mm;
    out->look_at = look_at;
    out->up = up;                                                                         // stored raw (cam.rs:56)
    out->right = rbrt_vec3{right.x, right.y, right.z};
    out->img_center_point = rbrt_vec3{position.x + k * la.x, position.y + k * la.y, position.z + k * la.z};
    out->mm_per_pix_hor = mm_per_pix_hor;
    out->mm_per_pix_vert = mm_per_pix_vert;
    return RBRT_OK;
}

// rbrt_transform_vertices (mesh.rs:102-112) lives in obj_loader.cpp, next to the loader that applies it.
