// adapter_test.cpp -- drives include/qr_gpu_mpc_adapter.hpp the way MPCStanceLegController does
// (Reset -> SetupProblem; SolveDenseMPC -> SolveMPCKernel + 12 x GetMPCSolution) on instances read from a
// binary dump of a golden fixture, and prints the forces for the Python test to compare.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "qr_gpu_mpc_adapter.hpp"

template <size_t N>
struct Vec {
    std::array<float, N> a;
    float* data() { return a.data(); }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[2];
    if (std::fread(hdr, sizeof(int), 2, f) != 2) return 2;
    const int B = hdr[0], h = hdr[1];
    float cfg[5 + 3 + 12 + 1];  // dt mu fmax mass | inertia | weights | alpha  (dt,mu,fmax,mass,pad)
    if (std::fread(cfg, sizeof(float), 21, f) != 21) return 2;
    Quadruped::SetupProblem(cfg[0], h, cfg[1], cfg[2], cfg[3], cfg + 5, cfg + 8, cfg[20]);
    for (int i = 0; i < B; ++i) {
        Vec<3> p, v, w, rpy;
        Vec<4> q;
        Vec<12> r;
        std::vector<float> traj(12 * h), gait(4 * h);
        bool ok = std::fread(p.data(), 4, 3, f) == 3 && std::fread(v.data(), 4, 3, f) == 3 &&
                  std::fread(q.data(), 4, 4, f) == 4 && std::fread(w.data(), 4, 3, f) == 3 &&
                  std::fread(r.data(), 4, 12, f) == 12 && std::fread(rpy.data(), 4, 3, f) == 3 &&
                  std::fread(traj.data(), 4, traj.size(), f) == traj.size() &&
                  std::fread(gait.data(), 4, gait.size(), f) == gait.size();
        if (!ok) return 2;
        Quadruped::SolveMPCKernel(p, v, q, w, r, rpy, traj.data(), gait.data());
        std::printf("F %d %d", i, Quadruped::GetMPCStatus());
        for (int leg = 0; leg < 4; ++leg)
            for (int axis = 0; axis < 3; ++axis) std::printf(" %.9g", Quadruped::GetMPCSolution(leg * 3 + axis));
        std::printf("\n");
    }
    std::fclose(f);
    return 0;
}
