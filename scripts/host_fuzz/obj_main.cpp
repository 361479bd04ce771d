#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/rbrt_gpu.h"
namespace rbrt { void set_error(const char* fmt, ...) { (void)fmt; } }
int main(int argc, char** argv) {
    int bad = 0;
    for (int i = 1; i < argc; ++i) {
        float* v = nullptr; uint64_t n = 0;
        rbrt_vec3 t{1, 2, 3}, r{0.1f, 0.2f, 0.3f};
        int rc = rbrt_mesh_load_obj(argv[i], t, r, 2.0f, &v, &n);
        if (rc != 0 && rc != RBRT_E_INVALID) ++bad;
        double s = 0; for (uint64_t k = 0; k < 9 * n; ++k) s += v[k];
        if (n > (1u << 30)) printf("%f", s);
        rbrt_mesh_free(v);
    }
    return bad ? 1 : 0;
}
