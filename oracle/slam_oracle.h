/* TEST INFRASTRUCTURE ONLY -- public surface of the CPU oracle (see slam_oracle.c header). */
#ifndef SLAMRS_ORACLE_H
#define SLAMRS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, theta; } so_pose; /* common/src/robot.rs:9-18 */
struct so_slam;

/* math.rs */
double so_prob_log_odds(double p);
double so_log_odds_probability(double l);
double so_angle_diff(double alpha, double beta);
/* robot.rs */
void so_odometry_new(float dl, float dr, float wheel, double out[4]);
so_pose so_odometry_sample(const double od[4], so_pose p, double z1, double z2);
double so_odometry_log_prob(const double od[4], so_pose a, so_pose b);
/* ray.rs / map.rs */
int64_t so_ray_cells(float x0, float y0, float x1, float y1, uint64_t size_x, uint64_t size_y, uint64_t extra,
                     int32_t* out_xy, int64_t cap);
int so_inverse_sensor_model(float distance, float measured_distance, int was_hit, float tolerance);
uint64_t so_grid_cells(float extent, float resolution);
/* slam.rs */
struct so_slam* so_create(float pos_x, float pos_y, float width, float height, float resolution,
                          uint64_t n_particles, int track_counts);
struct so_slam* so_create_ex(float pos_x, float pos_y, float width, float height, float resolution,
                             uint64_t n_particles, int track_counts, int sparse);
void so_destroy(struct so_slam* s);
void so_set_weight_override(struct so_slam* s, const double* raw);
void so_get_own_raw(const struct so_slam* s, double* out);
void so_set_threads(struct so_slam* s, int threads);
void so_set_adaptive_resampling(struct so_slam* s, double tau);
int so_resampled(const struct so_slam* s);
void so_set_dead_likelihood(struct so_slam* s, int on);
void so_set_trace(struct so_slam* s, int64_t particle, int64_t cap);
int so_update(struct so_slam* s, const double* angle, const double* dist, const uint8_t* valid, uint64_t nb,
              float dl, float dr, float wheel, const double* z, double u01);
int so_resample_fold(const double* raw, uint64_t n_particles, double u01, double* norm, double* cum, uint64_t* idx,
                     uint64_t* max_particle);
uint64_t so_n(const struct so_slam* s);
uint64_t so_grid_w(const struct so_slam* s);
uint64_t so_grid_h(const struct so_slam* s);
uint64_t so_max_particle(const struct so_slam* s);
void so_get_poses(const struct so_slam* s, float* out_xyt);
void so_set_poses(struct so_slam* s, const float* xyt);
void so_get_weights(const struct so_slam* s, double* norm, double* raw);
double so_number_of_effective_particles(const struct so_slam* s);
void so_get_indices(const struct so_slam* s, uint64_t* idx);
void so_get_odds(const struct so_slam* s, uint64_t particle, double* out);
int so_get_counts(const struct so_slam* s, uint64_t particle, uint16_t* n_free, uint16_t* n_occ);
so_pose so_estimated_pose(const struct so_slam* s);
void so_estimated_likelihood(const struct so_slam* s, double* out);
int64_t so_get_trace(const struct so_slam* s, int32_t* out, int64_t cap);
int so_clamped(const struct so_slam* s);
/* simulator restatement (synthetic input) */
uint64_t so_sim_scan(const float* segments, uint64_t n_seg, float px, float py, float ptheta, uint64_t n_beams,
                     float scanner_range, double* angle, double* dist, uint8_t* valid);
void so_sim_motion(float* px, float* py, float* ptheta, float sl, float sr, float wheel_base);
void so_libm_sincosf(const float* x, uint64_t n, float* s, float* c);
#ifdef __cplusplus
}
#endif
#endif
