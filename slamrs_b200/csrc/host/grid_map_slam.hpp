// C++ host-side mirror of the reference's grid SLAM interface, header-only, over the C ABI
// (include/slamrs_gpu.h). The reference is compiled Rust; no Rust toolchain exists in this
// image, so this C++ layer stands where the Rust wrapper crate would (INTEGRATION.md shows that
// crate). Names and argument meaning follow the Rust items:
//
//   Pose, Measurement, Observation, Odometry   slamrs/common/src/robot.rs:9-184
//   Probability                                slamrs/common/src/math.rs:8-47
//   GridData<T>                                slamrs/slam/src/grid/map.rs:181-264
//   GridMapSlamConfig, GridMapSlam             slamrs/slam/src/grid/slam.rs:13-97
//
// Error behaviour: the reference API is infallible (update() returns nothing), so failures
// surface as a thrown std::runtime_error carrying slamrs_gpu_last_error(); construction with
// zero particles throws like the reference's assert (particle.rs:16).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/slamrs_gpu.h"

namespace slamrs_host {

struct Pose {  // robot.rs:9-18
    float x = 0.f, y = 0.f, theta = 0.f;
};

struct Measurement {  // robot.rs:82-94
    double angle = 0.0, distance = 0.0, strength = 1.0;
    bool valid = true;
};

struct Observation {  // robot.rs:51-54
    std::size_t id = 0;
    std::vector<Measurement> measurements;
};

struct Odometry {  // robot.rs:115-129; the two Normals are derived inside the library (robot.rs:132-150)
    float distance_left, distance_right, wheel_distance;
    static Odometry create(float l, float r, float wheel) { return Odometry{l, r, wheel}; }  // Odometry::new
};

struct Probability {  // math.rs:8
    double v;
    double value() const { return v; }
};

template <typename T>
struct GridData {  // map.rs:181-187; index = row * size_y + column (map.rs:201-204)
    std::size_t size_x = 0, size_y = 0;
    std::vector<T> data;
    const T& get(std::size_t column, std::size_t row) const { return data[row * size_y + column]; }
};

struct GridMapSlamConfig {  // slam.rs:18-25
    float position[2] = {-2.f, -2.f};
    float width = 4.f, height = 4.f, resolution = 0.02f;
    std::size_t n_particles = 10;
};

struct GpuPlacement {  // what the YAML cannot carry
    int device = -1;
    uint32_t rank = 0, world_size = 1;
    uint64_t seed = 0x5EED5A11ull;
    uint32_t rng_mode = SLAMRS_RNG_SHARED_STREAM;
    uint32_t spare_slots = 0;
    uint32_t flags = 0;  // enum slamrs_flags
    uint32_t slot_cells = 0;  // 0 = whole-grid slots, power of two >= 256 = windowed slots
    uint8_t nccl_id[SLAMRS_NCCL_ID_BYTES] = {0};
};

class GridMapSlam {
public:
    // GridMapSlam::new, slam.rs:28-43
    explicit GridMapSlam(const GridMapSlamConfig& config, const GpuPlacement& pl = GpuPlacement()) : config_(config) {
        if (config.n_particles == 0) throw std::runtime_error("Must have at least one particle");
        slamrs_gpu_config c{};
        c.struct_size = sizeof(c);
        c.abi_version = SLAMRS_GPU_ABI_VERSION;
        c.pos_x = config.position[0];
        c.pos_y = config.position[1];
        c.resolution = config.resolution;
        check(slamrs_gpu_grid_cells(config.width, config.resolution, &c.grid_w), nullptr);
        check(slamrs_gpu_grid_cells(config.height, config.resolution, &c.grid_h), nullptr);
        c.n_particles = config.n_particles;
        c.seed = pl.seed;
        c.rng_mode = pl.rng_mode;
        c.device = pl.device;
        c.rank = pl.rank;
        c.world_size = pl.world_size;
        c.spare_slots = pl.spare_slots;
        c.flags = pl.flags;
        c.slot_cells = pl.slot_cells;
        for (int i = 0; i < SLAMRS_NCCL_ID_BYTES; ++i) c.nccl_id[i] = pl.nccl_id[i];
        grid_w_ = c.grid_w;
        grid_h_ = c.grid_h;
        check(slamrs_gpu_create(&c, &h_), nullptr);
    }
    ~GridMapSlam() { slamrs_gpu_destroy(h_); }  // Drop
    GridMapSlam(const GridMapSlam&) = delete;
    GridMapSlam& operator=(const GridMapSlam&) = delete;
    GridMapSlam(GridMapSlam&& o) noexcept : h_(o.h_), config_(o.config_), grid_w_(o.grid_w_), grid_h_(o.grid_h_) { o.h_ = nullptr; }

    // GridMapSlam::update(&mut self, z: &Observation, u: Odometry), slam.rs:46-75
    void update(const Observation& z, Odometry u, const double* z_draws = nullptr, const double* resample_u = nullptr) {
        const std::size_t n = z.measurements.size();
        angle_.resize(n); dist_.resize(n); valid_.resize(n);
        for (std::size_t i = 0; i < n; ++i) {
            angle_[i] = static_cast<float>(z.measurements[i].angle);     // `m.angle as f32`, map.rs:76
            dist_[i] = static_cast<float>(z.measurements[i].distance);   // `m.distance as f32`
            valid_[i] = z.measurements[i].valid ? 1 : 0;
        }
        check(slamrs_gpu_update(h_, angle_.data(), dist_.data(), valid_.data(), static_cast<uint32_t>(n), u.distance_left,
                                u.distance_right, u.wheel_distance, z_draws, resample_u), h_);
    }

    Pose estimated_pose() const {  // slam.rs:77-81
        float xyt[3];
        check(slamrs_gpu_pose(h_, xyt), h_);
        return Pose{xyt[0], xyt[1], xyt[2]};
    }

    GridData<Probability> estimated_likelihood() const {  // slam.rs:83-88
        GridData<Probability> g;
        g.size_x = grid_w_;
        g.size_y = grid_h_;
        g.data.resize(static_cast<std::size_t>(grid_w_) * grid_h_);
        static_assert(sizeof(Probability) == sizeof(double), "Probability is a transparent f64 newtype");
        check(slamrs_gpu_map_probability(h_, reinterpret_cast<double*>(g.data.data())), h_);
        return g;
    }

    // Pipelined form for a node that publishes map t while step t+1 runs: the copy into `out` (size_x * size_y
    // cells of page-locked host memory) overlaps whatever is issued next; `out` is complete after map_wait().
    void estimated_likelihood_async(Probability* out) const {
        check(slamrs_gpu_map_probability_async(h_, reinterpret_cast<double*>(out)), h_);
    }
    void map_wait() const { check(slamrs_gpu_map_wait(h_), h_); }

    // The informed part of the same map as f32 (what the visualizer converts every cell to,
    // visualize.rs:247): window = {x0, y0, x1, y1} in cells, values row-major; every cell outside
    // the window is exactly 0.5. Cuts the per-scan device-to-host copy from 8 B/cell of the whole grid.
    struct MapWindow {
        int32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        std::vector<float> data;
    };
    MapWindow estimated_likelihood_window() const {
        MapWindow w;
        int32_t e[4];
        check(slamrs_gpu_map_extent(h_, e), h_);
        w.x0 = e[0]; w.y0 = e[1]; w.x1 = e[2]; w.y1 = e[3];
        w.data.resize(static_cast<std::size_t>(w.x1 - w.x0) * static_cast<std::size_t>(w.y1 - w.y0));
        if (!w.data.empty()) check(slamrs_gpu_map_window(h_, SLAMRS_MAP_F32, w.x0, w.y0, w.x1, w.y1, w.data.data()), h_);
        return w;
    }

    // ParticleFilter::number_of_effective_particles, particle.rs:59-65 (normalised weights of the last update)
    double number_of_effective_particles() const {
        double v = 0.0;
        check(slamrs_gpu_effective_particles(h_, &v), h_);
        return v;
    }

    std::pair<float, float> map_position() const { return {config_.position[0], config_.position[1]}; }  // slam.rs:90-96

    slamrs_gpu_handle* raw() const { return h_; }

private:
    static void check(int rc, slamrs_gpu_handle* h) {
        if (rc != SLAMRS_OK)
            throw std::runtime_error("slamrs_gpu error " + std::to_string(rc) + ": " + slamrs_gpu_last_error(h));
    }
    slamrs_gpu_handle* h_ = nullptr;
    GridMapSlamConfig config_;
    uint32_t grid_w_ = 0, grid_h_ = 0;
    std::vector<float> angle_, dist_;
    std::vector<uint8_t> valid_;
};

}  // namespace slamrs_host
