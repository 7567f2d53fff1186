// lrm_cuda — file-protocol driver: the reference's `cuda` executable (several_leg.cpp:17-224) over
// the C ABI of liblrm_b200.so.
//
//   dist_input_tx.bin, dist_input_ty.bin, dist_input_tz.bin   float32 SoA planes written by
//                                                             before.py:96-99
//   -> out_reachability.bin                                   one byte per point (saveArrayToFile<bool>)
//   -> out_dist_xx.bin, out_dist_xy.bin, out_dist_xz.bin      float32 planes read by after.py:120-151
//
// Same files, same order of operations (reachability pass, then distance pass, each printing the
// kernel-only milliseconds and ns per point like several_leg.cpp:151-155,189-193), same default leg
// (settings.h:58 RobotNumb = 1 -> get_M2_leg(0)).  What the reference fixes at compile time
// (settings.h) is a flag here:  lrm_cuda [--dir D] [--robot 0|1] [--azimuth A] [--fused]
// There is no CPU mode: without a CUDA device the program reports the library's error and exits 1
// (the reference's CUDA_CHECK_ERROR convention, cross_compiled.cu:12-20).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lrm_c.h"

namespace {

// readArrayFromFile<float>, math_util.cpp:63-87: length = file size / 4
bool read_plane(const std::string& path, std::vector<float>* out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        std::fprintf(stderr, "Error opening file: %s\n", path.c_str());
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out->resize((size_t)bytes / sizeof(float));
    const size_t got = out->empty() ? 0 : std::fread(out->data(), sizeof(float), out->size(), f);
    std::fclose(f);
    return got == out->size();
}

// saveArrayToFile, math_util.cpp:46-54
template <typename T>
bool save_array(const std::string& path, const T* data, size_t n) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        std::fprintf(stderr, "error saving file %s\n", path.c_str());
        return false;
    }
    const size_t put = n ? std::fwrite(data, sizeof(T), n, f) : 0;
    std::fclose(f);
    return put == n;
}

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "CUDA error in %s: %s\n", what, lrm_last_error());
    std::exit(EXIT_FAILURE);
}

}  // namespace

int main(int argc, char** argv) {
    std::string dir = ".";
    int robot = 1;
    float azimuth = 0.f;
    bool fused = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--dir" && i + 1 < argc) dir = argv[++i];
        else if (a == "--robot" && i + 1 < argc) robot = std::atoi(argv[++i]);
        else if (a == "--azimuth" && i + 1 < argc) azimuth = (float)std::atof(argv[++i]);
        else if (a == "--fused") fused = true;
        else {
            std::fprintf(stderr, "usage: lrm_cuda [--dir D] [--robot 0|1] [--azimuth A] [--fused]\n");
            return 2;
        }
    }
    lrm_leg_t leg;
    if (lrm_default_leg(robot, azimuth, &leg) != LRM_OK) die("lrm_default_leg");

    std::vector<float> x, y, z;
    if (!read_plane(dir + "/dist_input_tx.bin", &x) || !read_plane(dir + "/dist_input_ty.bin", &y) ||
        !read_plane(dir + "/dist_input_tz.bin", &z))
        return EXIT_FAILURE;
    if (x.size() != y.size() || x.size() != z.size()) {
        std::fprintf(stderr, "input planes differ in length\n");
        return EXIT_FAILURE;
    }
    const size_t n = x.size();
    // threeArrays2float3Arr, math_util.cpp:92-104
    std::vector<float> xyz(3 * n);
    for (size_t i = 0; i < n; i++) xyz[3 * i] = x[i], xyz[3 * i + 1] = y[i], xyz[3 * i + 2] = z[i];

    std::vector<uint8_t> reach(n);
    std::vector<float> vec(3 * n);
    float ms = 0.f;
    if (fused) {
        // one pass for both outputs (no counterpart in the reference driver)
        if (lrm_reach_dist(xyz.data(), n, &leg, nullptr, reach.data(), vec.data(), 0, nullptr, &ms) != LRM_OK)
            die("lrm_reach_dist");
        std::printf("Cuda reachability + distance took %g milliseconds to finish.\n", ms);
        std::printf("That's %g ns per point (total: %zu)\n", n ? (double)ms / (double)n * 1e6 : 0.0, n);
    } else {
        if (lrm_reach(xyz.data(), n, &leg, nullptr, reach.data(), 0, nullptr, &ms) != LRM_OK) die("lrm_reach");
        std::printf("Cuda reachability took %g milliseconds to finish.\n", ms);
        std::printf("That's %g ns per point (total: %zu)\n", n ? (double)ms / (double)n * 1e6 : 0.0, n);
        if (lrm_dist(xyz.data(), n, &leg, nullptr, vec.data(), nullptr, 0, nullptr, &ms) != LRM_OK) die("lrm_dist");
        std::printf("Cuda distance took %g milliseconds to finish.\n", ms);
        std::printf("That's %g ns per point (total: %zu)\n", n ? (double)ms / (double)n * 1e6 : 0.0, n);
    }
    if (!save_array(dir + "/out_reachability.bin", reach.data(), n)) return EXIT_FAILURE;
    const char* names[3] = {"/out_dist_xx.bin", "/out_dist_xy.bin", "/out_dist_xz.bin"};
    std::vector<float> plane(n);
    for (int k = 0; k < 3; k++) {
        for (size_t i = 0; i < n; i++) plane[i] = vec[3 * i + k];
        if (!save_array(dir + names[k], plane.data(), n)) return EXIT_FAILURE;
    }
    return 0;
}
