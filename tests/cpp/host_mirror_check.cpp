// Exercises the C++ host mirror (slamrs_b200/csrc/host/grid_map_slam.hpp) the way
// GridMapSlamNode::update (slamrs/slam/src/grid/node.rs:47-60) drives the reference:
// new -> [update -> estimated_pose -> estimated_likelihood]* -> drop, twice (app.rs:121-134).
// Usage: host_mirror_check <scan.txt>   (lines: angle distance valid), prints a digest.
// With --no-gpu it only checks the error path (no CUDA device => exception, no fallback).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../slamrs_b200/csrc/host/grid_map_slam.hpp"

using namespace slamrs_host;

int main(int argc, char** argv) {
    if (argc >= 2 && std::strcmp(argv[1], "--no-gpu") == 0) {
        try {
            GridMapSlamConfig cfg;
            GridMapSlam slam(cfg);
            std::printf("UNEXPECTED: created a filter without a GPU\n");
            return 1;
        } catch (const std::exception& e) {
            std::printf("expected error: %s\n", e.what());
            return std::strstr(e.what(), "no CPU fallback") ? 0 : 2;
        }
    }
    try {
        GridMapSlamConfig zero;
        zero.n_particles = 0;
        GridMapSlam bad(zero);
        return 3;
    } catch (const std::exception&) {
    }
    Observation obs;
    std::ifstream in(argv[1]);
    double a, d;
    int v;
    while (in >> a >> d >> v) obs.measurements.push_back(Measurement{a, d, 1.0, v != 0});
    for (int cycle = 0; cycle < 2; ++cycle) {
        GridMapSlamConfig cfg;  // the shipped preset: 4x4 m at 2 cm, 10 particles
        GridMapSlam slam(cfg);
        for (int s = 0; s < 3; ++s) {
            slam.update(obs, Odometry::create(0.08f, 0.10f, 0.1f));
            const Pose p = slam.estimated_pose();
            const GridData<Probability> m = slam.estimated_likelihood();
            double occ = 0, fre = 0;
            for (const Probability& c : m.data) { occ += c.value() > 0.5; fre += c.value() < 0.5; }
            std::printf("cycle %d step %d pose %.9g %.9g %.9g cells %zux%zu occupied %.0f free %.0f\n", cycle, s, p.x, p.y,
                        p.theta, m.size_x, m.size_y, occ, fre);
        }
        const auto pos = slam.map_position();
        if (pos.first != -2.f || pos.second != -2.f) return 4;
    }
    return 0;
}
