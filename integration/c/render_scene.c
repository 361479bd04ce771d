/* render_scene.c — the drop-in boundary from plain C99: what `main.rs:70-91` does, over include/rbrt_gpu.h.
 *
 *   cc -std=c99 -I../../include render_scene.c -L../../rbrt_b200 -lrbrt_gpu -Wl,-rpath,$PWD/../../rbrt_b200 -o render_scene
 *   ./render_scene [mesh.obj] out.ppm
 *
 * Camera::new(position, look_at, up, HEIGHT, WIDTH, focal) (cam.rs:22-29) -> the spheres of scenes/example_scene.yaml:33-75 (+ an .obj loaded
 * and transformed like load_mesh_vertices_from_file, mesh.rs:78-121) -> create_scene_from_scene_blueprint -> render_scene -> a binary PPM.
 * Exit codes: 0 ok, 3 no usable GPU (the library has no CPU fallback), 1 anything else. */
#include <stdio.h>
#include <string.h>

#include "rbrt_gpu.h"

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rbrt_last_error());
    return 1;
}

int main(int argc, char** argv) {
    const char* obj = argc > 2 ? argv[1] : NULL;
    const char* out = argc > 2 ? argv[2] : (argc > 1 ? argv[1] : "out.ppm");
    const uint32_t width = 256, height = 192, samples = 8;

    rbrt_vec3 position = {0.0f, 5.0f, 4.0f}, look_at = {0.0f, -0.1f, -1.0f}, up = {0.0f, 1.0f, -0.4f};   /* example_scene.yaml:2-15 */
    rbrt_camera cam;
    if (rbrt_camera_new(position, look_at, up, height, width, 28.0f, &cam) != RBRT_OK) return fail("rbrt_camera_new");   /* height BEFORE width */

    rbrt_sphere_desc spheres[4];
    memset(spheres, 0, sizeof(spheres));
    {
        const float c[4][4] = {{0.0f, -1000.0f, -5.0f, 1000.0f}, {-5.0f, 1.5f, -9.0f, 1.5f}, {-2.5f, 2.9f, -15.0f, 3.0f}, {1.5f, 1.25f, -9.0f, 1.5f}};
        const rbrt_material m[4] = {{RBRT_MAT_LAMBERTIAN, {0.02f, 0.2f, 0.1f}, 0.0f}, {RBRT_MAT_LAMBERTIAN, {0.1f, 0.1f, 0.9f}, 0.0f},
                                    {RBRT_MAT_METAL, {0.8f, 0.8f, 0.8f}, 0.005f}, {RBRT_MAT_DIELECTRIC, {0.0f, 0.0f, 0.0f}, 1.8f}};
        int i;
        for (i = 0; i < 4; ++i) {
            spheres[i].center.x = c[i][0]; spheres[i].center.y = c[i][1]; spheres[i].center.z = c[i][2]; spheres[i].radius = c[i][3];
            spheres[i].material = m[i];
        }
    }

    rbrt_mesh_desc mesh;
    uint32_t n_meshes = 0;
    float* soup = NULL;
    memset(&mesh, 0, sizeof(mesh));
    if (obj) {                                                     /* example_scene.yaml:17-28: scale 45, translation (5, -1.8, -12.5), glass 0.2 */
        rbrt_vec3 translation = {5.0f, -1.8f, -12.5f}, rotation = {0.0f, 0.0f, 0.0f};
        uint64_t n = 0;
        if (rbrt_mesh_load_obj(obj, translation, rotation, 45.0f, &soup, &n) != RBRT_OK) return fail("rbrt_mesh_load_obj");   /* the reference panics here */
        printf("Successfully loaded %llu triangles from file %s!\n", (unsigned long long)n, obj);
        mesh.tri_vertices = soup; mesh.num_triangles = n;
        mesh.material.kind = RBRT_MAT_DIELECTRIC; mesh.material.param = 0.2f;
        n_meshes = 1;
    }

    {
        rbrt_scene* scene = NULL;
        static uint8_t rgb[192 * 256 * 3];
        rbrt_render_opts opts;
        rbrt_stats stats;
        FILE* f;
        int rc = rbrt_gpu_init(0);
        if (rc != RBRT_OK) { fail("rbrt_gpu_init"); rbrt_mesh_free(soup); return rc == RBRT_E_NODEVICE ? 3 : 1; }
        if (rbrt_gpu_scene_create(spheres, 4, &mesh, n_meshes, NULL, &scene) != RBRT_OK) { rbrt_mesh_free(soup); return fail("rbrt_gpu_scene_create"); }
        rbrt_mesh_free(soup);                                      /* the arrays have been read when scene_create returns */
        memset(&opts, 0, sizeof(opts));
        opts.seed = 7;
        if (rbrt_gpu_render(scene, &cam, samples, &opts, rgb, &stats) != RBRT_OK) { rbrt_gpu_scene_destroy(scene); return fail("rbrt_gpu_render"); }
        rbrt_gpu_scene_destroy(scene);
        f = fopen(out, "wb");
        if (!f) { fprintf(stderr, "Unable to save target img to %s!\n", out); return 1; }
        fprintf(f, "P6\n%u %u\n255\n", width, height);
        fwrite(rgb, 1, sizeof(rgb), f);
        fclose(f);
        printf("%llu rays in %.2f ms on the device\n", (unsigned long long)stats.rays, stats.ms_device);
    }
    return 0;
}
