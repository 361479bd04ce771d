// bvh_build.cuh — entry point of the GPU LBVH builder (bvh_build.cu).
#pragma once
#include "common.cuh"

namespace rbrt {

// upload_mesh copies a mesh's n_all x 9 floats (HOST pointer, world space, original order) into the process-wide build
// scratch and returns the exact AABB over ALL of them (aabbox.rs:62-88, computed on the device); build_mesh_bvh must
// follow it directly with d_raw = the pointer it returned and n = the triangles the reference actually tests.
cudaError_t upload_mesh(const float* h_tris, uint64_t n_all, float lo[3], float hi[3], const float** d_raw_out, cudaStream_t st);
void release_build_scratch();

// d_raw: n triangles x 9 floats on the device (world space, original order).
// Writes n triangle records (3 float4 each) in Morton order to d_tris, n unit normals in ORIGINAL
// order to d_normals and up to n-1 64-byte 4-wide nodes to d_nodes (tree_height = depth of the wide tree); qorg/qstep = the 16-bit grid the node boxes are
// quantised on.  lo/hi = exact mesh AABB.
// sah: run the tree-rotation pass during the refit (one-triangle leaves only).
cudaError_t build_mesh_bvh(const float* d_raw, uint32_t n, const float lo[3], const float hi[3], float pad, uint32_t leaf_size, bool sah,
                           float4* d_tris, float4* d_normals, float4* d_nodes, int32_t* root_ref, uint64_t* live_nodes,
                           int* tree_height, float qorg[3], float qstep[3], cudaStream_t st);

}  // namespace rbrt
