// run_ref_tests.cpp -- TEST INFRASTRUCTURE: runs the reference's OWN regression tests (tests/test_shading.cpp,
// tests/test_aux_channels.cpp, compiled from /root/reference where they lie) against a chosen renderer type, e.g.
// `--arch CUDA` = the backend of this repo behind Ray::CreateRenderer (oracle/cuda_binding/RendererCUDA.cpp).
// It plays the role of tests/main.cpp (whose test list is fixed and needs all 223 MB of textures): same globals, same
// test functions, same ref.tga gates, but the caller names the tests.  cwd must hold test_data/ (oracle/Makefile copies
// the meshes, the gold-scuffed texture set and every ref.tga there).
//   usage: test_ray_cuda --arch CUDA|REF|AVX2... [--list] [-j N] test_name... | --group untextured|complex5|all
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <string_view>
#include <vector>

#include "Ray.h"

#define T(name) void test_##name(const char *arch_list[], std::string_view preferred_device);
#define UNTEXTURED(X)                                                                                                  \
    X(oren_mat0) X(oren_mat1) X(oren_mat2) X(diff_mat0) X(diff_mat1) X(diff_mat2) X(sheen_mat0) X(sheen_mat1)          \
    X(sheen_mat2) X(sheen_mat3) X(glossy_mat0) X(glossy_mat1) X(glossy_mat2) X(spec_mat0) X(spec_mat1) X(spec_mat2)    \
    X(aniso_mat0) X(aniso_mat1) X(aniso_mat2) X(aniso_mat3) X(aniso_mat4) X(aniso_mat5) X(aniso_mat6) X(aniso_mat7)    \
    X(tint_mat0) X(tint_mat1) X(tint_mat2) X(plastic_mat0) X(plastic_mat1) X(plastic_mat2) X(metal_mat0) X(metal_mat1) \
    X(metal_mat2) X(emit_mat0) X(emit_mat1) X(coat_mat0) X(coat_mat1) X(coat_mat2) X(refr_mis0) X(refr_mis1)           \
    X(refr_mis2) X(refr_mat0) X(refr_mat1) X(refr_mat2) X(refr_mat3) X(trans_mat0) X(trans_mat1) X(trans_mat2)         \
    X(trans_mat3) X(trans_mat4) X(trans_mat5)
// textured (gold-scuffed set, one BC-compressed alpha map, one HDR environment map); complex_mat5_dir_light renders through a Filmic view transform; no procedural sky / cache
#define COMPLEX5(X)                                                                                                    \
    X(complex_mat5) X(complex_mat5_clipped) X(complex_mat5_adaptive) X(complex_mat5_regions) X(complex_mat5_nlm_filter) \
    X(complex_mat5_dof) X(complex_mat5_mesh_lights) X(complex_mat5_sphere_light) X(complex_mat5_inside_light)          \
    X(complex_mat5_spot_light) X(complex_mat5_dir_light) X(complex_mat5_hdri_light) X(two_sided_mat) X(aux_channels)       \
    X(ray_flags)
// the UNet denoiser (RendererBase::DenoiseImage(pass, region)) through InitUNetFilter with the tree's own weight set
#define UNET(X) X(complex_mat5_unet_filter)
UNTEXTURED(T)
COMPLEX5(T)
UNET(T)
#undef T

bool g_stop_on_fail = false;
std::atomic_bool g_tests_success{true};
std::atomic_bool g_log_contains_errors{false};
bool g_catch_flt_exceptions = false;
bool g_determine_sample_count = false;
bool g_minimal_output = true;
bool g_nohwrt = true;
bool g_nodx = true;
int g_validation_level = 0;

namespace {
struct Entry {
    const char *name;
    void (*fn)(const char *[], std::string_view);
    int group; // 0 untextured, 1 complex5
};
#define E0(name) {#name, test_##name, 0},
#define E1(name) {#name, test_##name, 1},
#define E2(name) {#name, test_##name, 2},
const Entry kTests[] = {UNTEXTURED(E0) COMPLEX5(E1) UNET(E2)};
} // namespace

int main(int argc, char *argv[]) {
    const char *arch[] = {"CUDA", nullptr};
    std::vector<std::string> names;
    std::string group;
    bool list = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--arch") && i + 1 < argc) {
            arch[0] = argv[++i];
        } else if (!strcmp(argv[i], "--group") && i + 1 < argc) {
            group = argv[++i];
        } else if (!strcmp(argv[i], "--list")) {
            list = true;
        } else {
            names.emplace_back(argv[i]);
        }
    }
    printf("Ray Version: %s, arch %s\n", Ray::Version(), arch[0]);
    int ran = 0, failed = 0;
    for (const Entry &e : kTests) {
        bool take = false;
        for (const std::string &n : names) {
            take |= (n == e.name);
        }
        take |= (group == "all" && e.group < 2) || (group == "untextured" && e.group == 0) ||
                (group == "complex5" && e.group == 1) || (group == "unet" && e.group == 2);
        if (list) {
            printf("%s %s\n", e.group == 0 ? "untextured" : (e.group == 1 ? "complex5  " : "unet      "), e.name);
            continue;
        }
        if (!take) {
            continue;
        }
        g_tests_success = true;
        g_log_contains_errors = false;
        e.fn(arch, std::string_view{});
        const bool ok = g_tests_success && !g_log_contains_errors;
        printf("RESULT %-28s %s%s\n", e.name, ok ? "PASS" : "FAIL", g_log_contains_errors ? " (ILog::Error was called)" : "");
        fflush(stdout);
        ++ran;
        failed += ok ? 0 : 1;
    }
    if (!list) {
        printf("SUMMARY arch %s: %d run, %d failed\n", arch[0], ran, failed);
    }
    return failed ? 1 : 0;
}
