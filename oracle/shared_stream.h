/* TEST INFRASTRUCTURE -- part of the CPU oracle, never linked into the product library.
 *
 * The "shared seeded stream" of the parity contract (BASELINE.json north_star: "motion-noise
 * and resample uniforms fed from one shared seeded stream"). The Rust reference draws from
 * rand::thread_rng() (slamrs/common/src/robot.rs:173) and rand::random::<f64>()
 * (slamrs/slam/src/grid/particle.rs:84), which are OS-seeded and cannot be reproduced; both
 * sides of every parity test therefore consume THIS stream instead.
 *
 * Specification (must match slamrs_b200/csrc/shared_stream.cuh bit for bit):
 *   bits   : Philox4x32-10 (Salmon et al., SC'11), key = (seed lo32, seed hi32),
 *            counter = (particle lo32, particle hi32 ... see below)
 *   motion : counter (particle, step_lo, step_hi, 0) -> words x0..x3
 *            u1 = (((x0<<32|x1) >> 11) + 1) * 2^-53   in (0,1]
 *            u2 =  ((x2<<32|x3) >> 11)      * 2^-53   in [0,1)
 *            r = sqrt(-2 * dlog(u1));  z1 = r * dcos2pi(u2);  z2 = r * dsin2pi(u2)
 *            z1 drives the centre distance, z2 the heading (robot.rs:175-176 draw order).
 *   resample uniform : counter (0, step_lo, step_hi, 1) -> U = ((x0<<32|x1) >> 11) * 2^-53
 *   dlog / dsin2pi / dcos2pi are fixed sequences of IEEE-754 binary64 +,-,*,/ (no FMA
 *   contraction, no libm), so CPU and GPU produce identical bits.
 */
#ifndef SLAMRS_ORACLE_SHARED_STREAM_H
#define SLAMRS_ORACLE_SHARED_STREAM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void ss_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double ss_dlog(double x);
void ss_dsincos2pi(double u, double* s, double* c);
/* two standard normals for (seed, step, particle) */
void ss_motion_normals(uint64_t seed, uint64_t step, uint64_t particle, double* z1, double* z2);
/* fills z[2*i], z[2*i+1] for particles first..first+count-1 */
void ss_fill_motion_normals(uint64_t seed, uint64_t step, uint64_t first, uint64_t count, double* z);
double ss_resample_uniform(uint64_t seed, uint64_t step);
/* uniform start poses over [x0,x1) x [y0,y1), heading in [-pi, pi) (global-localisation-style initialisation) */
void ss_uniform_pose(uint64_t seed, uint64_t particle, double x0, double y0, double x1, double y1, float* out_xyt);
void ss_fill_uniform_poses(uint64_t seed, uint64_t first, uint64_t count, double x0, double y0, double x1, double y1,
                           float* out_xyt);

#ifdef __cplusplus
}
#endif
#endif
