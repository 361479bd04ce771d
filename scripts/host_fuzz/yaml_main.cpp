#include <cstdio>
#include <fstream>
#include <sstream>
#include "../../rbrt_b200/csrc/host/scene_yaml.hpp"
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        std::ifstream f(argv[i], std::ios::binary); std::stringstream b; b << f.rdbuf();
        try { auto bp = scene_yaml::parse_scene(b.str()); if (bp.spheres.size() > 1000000) puts("big"); } catch (const scene_yaml::ParseError&) {}
    }
    return 0;
}
