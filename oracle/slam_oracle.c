/* ============================================================================================
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the grid particle-filter SLAM step.
 *
 * A line-by-line C restatement of the Rust reference's hot path (antbern/slamrs). It is used
 * ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * as the checker and as the timed CPU baseline. The product library (libslamrs_gpu.so) never
 * links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: **parity unpinned** for the grid module. The reference has no tests, fixtures
 * or golden vectors for slamrs/slam/src/grid/ and its RNG is OS-seeded, and no Rust toolchain
 * exists in this environment, so the reference itself cannot be run. What IS pinned:
 *   - common/src/math.rs:167-195 (log-odds round trip, zero_is_half, angle_diff cases) --
 *     reproduced in tests/test_oracle_reference_kats.py against this file;
 *   - slamrs/out.log:4 (222 valid beams for the first simulator scan) -- reproduced by the
 *     scan generator below;
 *   - Philox4x32-10 published known-answer vectors for the shared stream.
 * A second, independently written restatement (oracle/numpy_restatement.py) guards against a
 * shared misreading of the Rust source.
 *
 * Types follow the reference exactly: poses and all ray geometry are f32 (`float`), cell
 * log-odds, probabilities, log-weights, weights and resampling are f64 (`double`). Build with
 * -ffp-contract=off -fno-fast-math: rustc never contracts a*b+c. On x86-64 Linux Rust's
 * f32::cos/sin and f64::exp/ln call the platform libm, as this file does.
 *
 * Third-party arithmetic restated from the published algorithm (crates absent from
 * /root/reference, pinned in slamrs/Cargo.lock): statrs 0.18.0 Normal::pdf and
 * Normal::sample (= mean + std_dev * z), rand 0.8.5 random::<f64>() (53-bit uniform),
 * nalgebra 0.34.2 EuclideanNorm metric distance (sqrt of a left fold of squared diffs).
 * ============================================================================================ */
#include "slam_oracle.h"
#include "shared_stream.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- Rust cast semantics */
/* `f32 as usize`: saturating, NaN -> 0 */
static inline uint64_t f32_as_usize(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)v;
}
/* `f32 as isize`: saturating, NaN -> 0 */
static inline int64_t f32_as_isize(float v) {
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)v;
}
/* wrapping isize add (release-mode Rust) */
static inline int64_t wrapping_add_i64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }

/* ---------------------------------------------------------------- common/src/math.rs */
/* Probability::log_odds, math.rs:30-32 */
double so_prob_log_odds(double p) { return log(p / (1.0 - p)); }
/* LogOdds::probability, math.rs:135-137 */
double so_log_odds_probability(double l) { return 1.0 - 1.0 / (1.0 + exp(l)); }
/* angle_diff, math.rs:150-157 (Rust `%` on f64 is C fmod) */
double so_angle_diff(double alpha, double beta) {
    const double PI = 3.14159265358979323846264338327950288;
    double diff = fmod(beta - alpha + PI, PI * 2.0) - PI;
    if (diff < -PI) return diff + 2.0 * PI;
    return diff;
}

/* ---------------------------------------------------------------- common/src/robot.rs */
/* Odometry::new, robot.rs:132-150. out = {mean_c, std_c, mean_t, std_t} */
void so_odometry_new(float dl, float dr, float wheel, double out[4]) {
    double delta_center = (double)((dl + dr) / 2.0f);
    double delta_theta = (double)((dr - dl) / wheel);
    double delta_center_std = (0.01 + fabs(delta_center) * 0.05) / 2.0;
    /* f64::to_radians(5.0) = 5.0 * (PI / 180.0) */
    const double rads_per_deg = 3.14159265358979323846264338327950288 / 180.0;
    double delta_theta_std = 5.0 * rads_per_deg + 0.1 * fabs(delta_theta);
    out[0] = delta_center;
    out[1] = delta_center_std;
    out[2] = delta_theta;
    out[3] = delta_theta_std;
}

/* statrs 0.18.0 Normal::pdf: d = (x-mean)/std; exp(-0.5*d*d) / (SQRT_2PI*std) */
static inline double normal_pdf(double x, double mean, double std) {
    const double SQRT_2PI = 2.5066282746310005024157652848110452530069867406099;
    double d = (x - mean) / std;
    return exp(-0.5 * d * d) / (SQRT_2PI * std);
}

/* Odometry::sample, robot.rs:170-183; z1, z2 are the two standard-normal draws in draw order
 * (statrs sample = mean + std_dev * z). */
so_pose so_odometry_sample(const double od[4], so_pose p, double z1, double z2) {
    float center_distance = (float)(od[0] + od[1] * z1);
    float theta = p.theta + (float)(od[2] + od[3] * z2);
    so_pose n;
    n.theta = theta;
    n.x = p.x + cosf(theta) * center_distance;
    n.y = p.y + sinf(theta) * center_distance;
    return n;
}

/* Odometry::probabiliy_of, robot.rs:152-167 -> log probability (LogProbability product = sum) */
double so_odometry_log_prob(const double od[4], so_pose a, so_pose b) {
    float dx = a.x - b.x, dy = a.y - b.y;
    float center_distance = sqrtf(dx * dx + dy * dy); /* powi(2) == x*x */
    double angle_distance = so_angle_diff((double)a.theta, (double)b.theta);
    return log(normal_pdf((double)center_distance, od[0], od[1])) + log(normal_pdf(angle_distance, od[2], od[3]));
}

/* ---------------------------------------------------------------- slam/src/grid/ray.rs */
/* GridRayIterator::new (ray.rs:21-77) + next (ray.rs:83-110). Calls visit(x, y, ctx) per cell. */
typedef void (*visit_fn)(int64_t x, int64_t y, void* ctx);

static void ray_walk(float x0, float y0, float x1, float y1, uint64_t size_x, uint64_t size_y,
                     uint64_t additional_steps, visit_fn visit, void* ctx) {
    float delta_x = fabsf(x1 - x0), delta_y = fabsf(y1 - y0);
    int64_t x = f32_as_isize(floorf(x0));
    int64_t y = f32_as_isize(floorf(y0));
    int64_t n = 1 + (int64_t)additional_steps;
    int64_t x_inc, y_inc;
    float error;
    if (delta_x == 0.0f) {
        x_inc = 0;
        error = INFINITY;
    } else if (x1 > x0) {
        x_inc = 1;
        n = wrapping_add_i64(n, f32_as_isize(floorf(x1) - (float)x));
        error = (floorf(x0) + 1.0f - x0) * delta_y;
    } else {
        x_inc = -1;
        n = wrapping_add_i64(n, (int64_t)((uint64_t)x - (uint64_t)f32_as_isize(floorf(x1))));
        error = (x0 - floorf(x0)) * delta_y;
    }
    if (delta_y == 0.0f) {
        y_inc = 0;
        error -= INFINITY;
    } else if (y1 > y0) {
        y_inc = 1;
        n = wrapping_add_i64(n, (int64_t)((uint64_t)f32_as_isize(floorf(y1)) - (uint64_t)y));
        error -= (floorf(y0) + 1.0f - y0) * delta_x;
    } else {
        y_inc = -1;
        n = wrapping_add_i64(n, (int64_t)((uint64_t)y - (uint64_t)f32_as_isize(floorf(y1))));
        error -= (y0 - floorf(y0)) * delta_x;
    }
    uint64_t remaining = (uint64_t)n; /* `n as usize` */
    for (;;) {
        int one_more = remaining > 0 && !(x < 0 || x >= (int64_t)size_x || y < 0 || y >= (int64_t)size_y);
        if (!one_more) break;
        visit(x, y, ctx);
        if (error > 0.0f) {
            y += y_inc;
            error -= delta_x;
        } else {
            x += x_inc;
            error += delta_y;
        }
        remaining -= 1;
    }
}

typedef struct {
    int32_t* out;
    int64_t cap, count;
} collect_ctx;
static void collect_visit(int64_t x, int64_t y, void* c) {
    collect_ctx* cc = (collect_ctx*)c;
    if (cc->count < cc->cap) {
        cc->out[2 * cc->count] = (int32_t)x;
        cc->out[2 * cc->count + 1] = (int32_t)y;
    }
    cc->count++;
}
/* test hook: list of (x,y) cells visited; returns the number of cells (may exceed cap) */
int64_t so_ray_cells(float x0, float y0, float x1, float y1, uint64_t size_x, uint64_t size_y, uint64_t extra,
                     int32_t* out_xy, int64_t cap) {
    collect_ctx c = {out_xy, cap, 0};
    ray_walk(x0, y0, x1, y1, size_x, size_y, extra, collect_visit, &c);
    return c.count;
}

/* ---------------------------------------------------------------- slam/src/grid/map.rs */
/* inverse_sensor_model, map.rs:148-172. returns 0 = P_PRIOR, 1 = P_FREE, 2 = P_OCCUPPIED */
int so_inverse_sensor_model(float distance, float measured_distance, int was_hit, float tolerance) {
    if (!was_hit) {
        if (distance < measured_distance) return 1;
        return 0;
    }
    if (distance < measured_distance - tolerance / 2.0f) return 1;
    if (distance > measured_distance + tolerance / 2.0f) return 0;
    return 2;
}

/* Sparse storage (tests at BASELINE.json's full populations only; so_create_ex(..., sparse = 1)): the grid
 * is cut into SO_TILE x SO_TILE tiles that exist only once a cell in them has been written, and
 * `clone()` shares tiles by reference count until one side writes (copy on write). Cell values, the
 * order of the additions and every index computation are the dense path's; only where a cell lives
 * changes. 8,192 particles with 1024^2 f64 grids need 64 GiB dense (twice that during resample) and a
 * few GiB this way. The dense layout stays the default and is what the CPU baseline times. */
#define SO_TILE 32
struct so_tile {
    int refs;
    double odds[SO_TILE * SO_TILE];
    uint16_t n_free[SO_TILE * SO_TILE];
    uint16_t n_occ[SO_TILE * SO_TILE];
};

struct so_map {
    float pos_x, pos_y, res;
    uint64_t gw, gh;
    double* odds;      /* gw*gh f64 log-odds, index = row*gh + column (map.rs:201-204) */
    uint16_t* n_free;  /* optional exact counters for integer parity checks */
    uint16_t* n_occ;
    struct so_tile** tiles; /* sparse storage: tiles_x * tiles_y pointers, NULL = every cell still at the prior */
    uint64_t tiles_x, tiles_y;
    double prior;
};

struct so_slam {
    uint64_t n;
    float pos_x, pos_y, res;
    uint64_t gw, gh;
    so_pose* pose;
    struct so_map* map;
    double* weight; /* normalised weights of the last update, pre-resample order */
    double* raw_weight;
    uint64_t* last_idx;
    uint64_t max_particle;
    int track_counts;
    int sparse;            /* tile storage with copy-on-write clones (so_create_ex) */
    const double* weight_override; /* next update resamples on these raw weights (so_set_weight_override) */
    double* own_raw;       /* the raw weights this oracle computed in the last update */
    double adaptive_tau;   /* > 0: resample only when N_eff < tau * N (extension, see so_set_adaptive_resampling) */
    double* carry;         /* weights carried over a step that did not resample; NULL after a resampling */
    int resampled;         /* did the last update resample? */
    int run_dead_likelihood;
    int clamped; /* resample index ran past N-1 (reference would panic) */
    int threads;
    double l_free, l_occ, l_prior;
    /* optional trace of one particle's visited cells in the last update */
    int64_t trace_particle;
    int32_t* trace;   /* triples (x, y, kind) */
    int64_t trace_cap, trace_count;
};

static inline size_t cell_index(const struct so_map* m, uint64_t column, uint64_t row) { return row * m->gh + column; }

/* sparse storage: tile of a cell and the cell's place inside it */
static inline size_t tile_of(const struct so_map* m, uint64_t column, uint64_t row) { return (row / SO_TILE) * m->tiles_x + column / SO_TILE; }
static inline size_t in_tile(uint64_t column, uint64_t row) { return (row % SO_TILE) * SO_TILE + column % SO_TILE; }
static inline double map_read_odds(const struct so_map* m, uint64_t column, uint64_t row) {
    if (!m->tiles) return m->odds[cell_index(m, column, row)];
    const struct so_tile* t = m->tiles[tile_of(m, column, row)];
    return t ? t->odds[in_tile(column, row)] : m->prior;
}
static void tile_release(struct so_tile* t) {
    if (t && __atomic_sub_fetch(&t->refs, 1, __ATOMIC_ACQ_REL) == 0) free(t);
}
/* the tile of a cell, private to this map and ready to be written */
static struct so_tile* map_write_tile(struct so_map* m, uint64_t column, uint64_t row) {
    struct so_tile** slot = &m->tiles[tile_of(m, column, row)];
    struct so_tile* t = *slot;
    if (t && __atomic_load_n(&t->refs, __ATOMIC_ACQUIRE) == 1) return t;
    struct so_tile* fresh = (struct so_tile*)malloc(sizeof(struct so_tile));
    if (!fresh) abort();
    if (t) {
        memcpy(fresh, t, sizeof(*fresh));       /* copy first, release after: the last holder frees */
    } else {
        for (int c = 0; c < SO_TILE * SO_TILE; ++c) fresh->odds[c] = m->prior;
        memset(fresh->n_free, 0, sizeof(fresh->n_free));
        memset(fresh->n_occ, 0, sizeof(fresh->n_occ));
    }
    fresh->refs = 1;
    *slot = fresh;
    tile_release(t);
    return fresh;
}

/* Map::world_to_grid, map.rs:60-62 */
static inline void world_to_grid(const struct so_map* m, float wx, float wy, float* gx, float* gy) {
    *gx = (wx - m->pos_x) / m->res;
    *gy = (wy - m->pos_y) / m->res;
}
/* Map::is_valid, map.rs:64-69 */
static inline int is_valid(const struct so_map* m, float gx, float gy) {
    return !((gx < 0.0f) || (gy < 0.0f) || (f32_as_usize(gx) >= m->gw) || (f32_as_usize(gy) >= m->gh));
}

/* Map::probability_of, map.rs:113-145 -> LogProbability value */
static double map_log_probability_of(const struct so_map* m, const double* angle, const double* dist,
                                     const uint8_t* valid, uint64_t nb, so_pose pose) {
    const double Z_HIT = 0.9, SENSOR_MAXDIST = 1.0;
    double product = log(1.0);
    for (uint64_t i = 0; i < nb; ++i) {
        if (!valid[i]) continue;
        float a = pose.theta + (float)angle[i];
        float ex = pose.x + cosf(a) * (float)dist[i];
        float ey = pose.y + sinf(a) * (float)dist[i];
        float gx, gy;
        world_to_grid(m, ex, ey, &gx, &gy);
        if (is_valid(m, gx, gy)) {
            double odds = map_read_odds(m, f32_as_usize(gx), f32_as_usize(gy));
            double p = so_log_odds_probability(odds);
            if (p == 0.5) {
                product += log(1.0 / SENSOR_MAXDIST);
            } else {
                product += log(Z_HIT * p + (1.0 - Z_HIT) * 1.0 / SENSOR_MAXDIST);
            }
        }
    }
    return product;
}

typedef struct {
    struct so_map* m;
    struct so_slam* s;
    float sx, sy, measured;
    int was_hit;
    int tracing;
} integrate_ctx;

static void integrate_visit(int64_t x, int64_t y, void* c) {
    integrate_ctx* ic = (integrate_ctx*)c;
    struct so_map* m = ic->m;
    float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
    /* nalgebra EuclideanNorm::metric_distance: fold from 0 of (a-b)^2, then sqrt */
    float dxx = ic->sx - cx, dyy = ic->sy - cy;
    float acc = 0.0f;
    acc = acc + dxx * dxx;
    acc = acc + dyy * dyy;
    float distance = sqrtf(acc);
    int kind = so_inverse_sensor_model(distance, ic->measured, ic->was_hit, 2.0f);
    double inc = kind == 1 ? ic->s->l_free : (kind == 2 ? ic->s->l_occ : ic->s->l_prior);
    if (m->tiles) {
        struct so_tile* t = map_write_tile(m, (uint64_t)x, (uint64_t)y);
        size_t idx = in_tile((uint64_t)x, (uint64_t)y);
        t->odds[idx] += inc;
        if (kind == 1 && t->n_free[idx] != UINT16_MAX) t->n_free[idx]++;
        if (kind == 2 && t->n_occ[idx] != UINT16_MAX) t->n_occ[idx]++;
    } else {
        size_t idx = cell_index(m, (uint64_t)x, (uint64_t)y);
        m->odds[idx] += inc;
        if (m->n_free) {
            if (kind == 1 && m->n_free[idx] != UINT16_MAX) m->n_free[idx]++;
            if (kind == 2 && m->n_occ[idx] != UINT16_MAX) m->n_occ[idx]++;
        }
    }
    if (ic->tracing) {
        struct so_slam* s = ic->s;
        if (s->trace_count < s->trace_cap) {
            s->trace[3 * s->trace_count] = (int32_t)x;
            s->trace[3 * s->trace_count + 1] = (int32_t)y;
            s->trace[3 * s->trace_count + 2] = kind;
        }
        s->trace_count++;
    }
}

/* Map::integrate + apply_measurement, map.rs:71-106 */
static void map_integrate(struct so_slam* s, struct so_map* m, const double* angle, const double* dist,
                          const uint8_t* valid, uint64_t nb, so_pose pose, int tracing) {
    float sx, sy;
    world_to_grid(m, pose.x, pose.y, &sx, &sy);
    for (uint64_t i = 0; i < nb; ++i) {
        float a = pose.theta + (float)angle[i];
        float ex = pose.x + cosf(a) * (float)dist[i];
        float ey = pose.y + sinf(a) * (float)dist[i];
        float gx, gy;
        world_to_grid(m, ex, ey, &gx, &gy);
        integrate_ctx ic = {m, s, sx, sy, (float)dist[i] / m->res, valid[i] != 0, tracing};
        ray_walk(sx, sy, gx, gy, m->gw, m->gh, 2, integrate_visit, &ic);
    }
}

/* ---------------------------------------------------------------- lifecycle */
static int map_alloc(struct so_map* m, const struct so_slam* s) {
    m->pos_x = s->pos_x; m->pos_y = s->pos_y; m->res = s->res; m->gw = s->gw; m->gh = s->gh;
    m->tiles = NULL; m->prior = s->l_prior;
    if (s->sparse) {
        m->odds = NULL; m->n_free = m->n_occ = NULL;
        m->tiles_x = (s->gw + SO_TILE - 1) / SO_TILE; m->tiles_y = (s->gh + SO_TILE - 1) / SO_TILE;
        m->tiles = (struct so_tile**)calloc((size_t)(m->tiles_x * m->tiles_y), sizeof(struct so_tile*));
        return m->tiles ? 0 : -1;
    }
    size_t cells = (size_t)(s->gw * s->gh);
    m->odds = (double*)malloc(cells * sizeof(double));
    m->n_free = m->n_occ = NULL;
    if (!m->odds) return -1;
    if (s->track_counts) {
        m->n_free = (uint16_t*)calloc(cells, 2);
        m->n_occ = (uint16_t*)calloc(cells, 2);
        if (!m->n_free || !m->n_occ) return -1;
    }
    return 0;
}
static void map_free(struct so_map* m) {
    if (m->tiles) {
        for (size_t t = 0; t < (size_t)(m->tiles_x * m->tiles_y); ++t) tile_release(m->tiles[t]);
        free(m->tiles);
        m->tiles = NULL;
    }
    free(m->odds); free(m->n_free); free(m->n_occ);
    m->odds = NULL; m->n_free = m->n_occ = NULL;
}
static void map_copy(struct so_map* dst, const struct so_map* src) {
    if (src->tiles) {   /* clone(): share every tile until one side writes it */
        for (size_t t = 0; t < (size_t)(src->tiles_x * src->tiles_y); ++t) {
            struct so_tile* tile = src->tiles[t];
            if (tile) __atomic_add_fetch(&tile->refs, 1, __ATOMIC_ACQ_REL);
            dst->tiles[t] = tile;
        }
        return;
    }
    size_t cells = (size_t)(src->gw * src->gh);
    memcpy(dst->odds, src->odds, cells * sizeof(double));
    if (src->n_free) {
        memcpy(dst->n_free, src->n_free, cells * 2);
        memcpy(dst->n_occ, src->n_occ, cells * 2);
    }
}

/* Map::new grid size, map.rs:28-31: ceil(width / resolution) as usize in f32 */
uint64_t so_grid_cells(float extent, float resolution) { return f32_as_usize(ceilf(extent / resolution)); }

/* GridMapSlam::new, slam.rs:28-43 (+ Map::new map.rs:26-48, ParticleFilter::new particle.rs:15-28) */
struct so_slam* so_create(float pos_x, float pos_y, float width, float height, float resolution,
                          uint64_t n_particles, int track_counts) {
    return so_create_ex(pos_x, pos_y, width, height, resolution, n_particles, track_counts, 0);
}

struct so_slam* so_create_ex(float pos_x, float pos_y, float width, float height, float resolution,
                             uint64_t n_particles, int track_counts, int sparse) {
    if (n_particles == 0) return NULL; /* reference asserts */
    struct so_slam* s = (struct so_slam*)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->sparse = sparse;
    s->n = n_particles;
    s->pos_x = pos_x; s->pos_y = pos_y; s->res = resolution;
    s->gw = so_grid_cells(width, resolution);
    s->gh = so_grid_cells(height, resolution);
    s->track_counts = track_counts;
    s->threads = 1;
    s->trace_particle = -1;
    s->l_free = so_prob_log_odds(0.30);
    s->l_occ = so_prob_log_odds(0.9);
    s->l_prior = so_prob_log_odds(0.5);
    s->pose = (so_pose*)calloc(n_particles, sizeof(so_pose)); /* Pose::default() */
    s->map = (struct so_map*)calloc(n_particles, sizeof(struct so_map));
    s->weight = (double*)malloc(n_particles * sizeof(double));
    s->raw_weight = (double*)malloc(n_particles * sizeof(double));
    s->last_idx = (uint64_t*)calloc(n_particles, sizeof(uint64_t));
    s->own_raw = (double*)calloc(n_particles, sizeof(double));
    if (!s->pose || !s->map || !s->weight || !s->raw_weight || !s->last_idx || !s->own_raw) { so_destroy(s); return NULL; }
    size_t cells = (size_t)(s->gw * s->gh);
    double init = so_prob_log_odds(0.5);
    for (uint64_t i = 0; i < n_particles; ++i) {
        if (map_alloc(&s->map[i], s)) { so_destroy(s); return NULL; }
        if (!s->sparse) for (size_t c = 0; c < cells; ++c) s->map[i].odds[c] = init;
        s->weight[i] = 1.0 / (double)n_particles;
        s->raw_weight[i] = s->weight[i];
        s->last_idx[i] = i;
    }
    s->max_particle = 0;
    return s;
}

void so_destroy(struct so_slam* s) {
    if (!s) return;
    if (s->map) for (uint64_t i = 0; i < s->n; ++i) map_free(&s->map[i]);
    free(s->map); free(s->pose); free(s->weight); free(s->raw_weight); free(s->last_idx); free(s->trace); free(s->own_raw); free(s->carry);
    free(s);
}

/* extension: resample only when N_eff < tau * N; 0 restores the reference's behaviour */
void so_set_adaptive_resampling(struct so_slam* s, double tau) { s->adaptive_tau = tau; }
int so_resampled(const struct so_slam* s) { return s->resampled; }
void so_set_threads(struct so_slam* s, int threads) { s->threads = threads < 1 ? 1 : threads; }
void so_set_dead_likelihood(struct so_slam* s, int on) { s->run_dead_likelihood = on; }
void so_set_trace(struct so_slam* s, int64_t particle, int64_t cap) {
    free(s->trace);
    s->trace = NULL;
    s->trace_particle = particle;
    s->trace_cap = cap;
    s->trace_count = 0;
    if (particle >= 0 && cap > 0) s->trace = (int32_t*)malloc((size_t)cap * 3 * sizeof(int32_t));
}

static volatile double g_sink;

/* the closure body of GridMapSlam::update, slam.rs:51-72 */
static void particle_step(struct so_slam* s, uint64_t p, const double* angle, const double* dist, const uint8_t* valid,
                          uint64_t nb, const double od[4], const double* z) {
    so_pose initial = s->pose[p];
    so_pose np = so_odometry_sample(od, initial, z[2 * p], z[2 * p + 1]);
    struct so_map* m = &s->map[p];
    if (s->run_dead_likelihood && !s->sparse) {
        /* slam.rs:58 `let likelihood = map.likelihood();` -- result unused in the reference */
        size_t cells = (size_t)(m->gw * m->gh);
        double* tmp = (double*)malloc(cells * sizeof(double));
        for (size_t c = 0; c < cells; ++c) tmp[c] = so_log_odds_probability(m->odds[c]);
        g_sink = tmp[cells / 2];
        free(tmp);
    }
    double lw = map_log_probability_of(m, angle, dist, valid, nb, np) + so_odometry_log_prob(od, initial, np);
    map_integrate(s, m, angle, dist, valid, nb, np, (int64_t)p == s->trace_particle);
    s->pose[p] = np;
    s->raw_weight[p] = exp(lw); /* weight.prob().value(), slam.rs:71 */
}

/* f64::total_cmp */
static inline int total_cmp(double a, double b) {
    int64_t x, y;
    memcpy(&x, &a, 8); memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
    y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return (x > y) - (x < y);
}

/* ParticleFilter::update's tail and ParticleFilter::resample up to the index vector, on caller-supplied raw
 * weights: normalize_weights (particle.rs:49-56: sequential sum, divide), argmax by total_cmp with the
 * last maximum winning (particle.rs:40-46), systematic resampling (particle.rs:78-101). cum (optional)
 * receives the value of `c` after weight i has been added (c = weights[0] for i = 0). Returns 1 if the
 * index ran past N-1 (the reference would panic with an out-of-bounds index), else 0. */
int so_resample_fold(const double* raw, uint64_t n_particles, double u01, double* norm, double* cum, uint64_t* idx,
                     uint64_t* max_particle) {
    const int64_t n = (int64_t)n_particles;
    /* normalize_weights, particle.rs:49-56 */
    double sum = 0.0;
    for (int64_t p = 0; p < n; ++p) sum += raw[p];
    for (int64_t p = 0; p < n; ++p) norm[p] = raw[p] / sum;
    /* max_by(total_cmp): last maximum wins, particle.rs:40-46 */
    uint64_t best = 0;
    for (int64_t p = 1; p < n; ++p)
        if (total_cmp(norm[p], norm[best]) >= 0) best = (uint64_t)p;
    *max_particle = best;

    /* resample, particle.rs:78-105 */
    int clamped = 0;
    double num = (double)n_particles;
    double r = u01 * 1.0 / num;
    double c = norm[0];
    uint64_t i = 0;
    if (cum) cum[0] = c;
    for (uint64_t mm = 1; mm <= n_particles; ++mm) {
        double u = r + ((double)mm - 1.0) * 1.0 / num;
        while (u > c) {
            if (i + 1 >= n_particles) { clamped = 1; break; } /* reference: index-out-of-bounds panic */
            i += 1;
            c += norm[i];
            if (cum) cum[i] = c;
        }
        idx[mm - 1] = i;
    }
    if (cum) /* the loop stops adding once every threshold is met; the remaining prefixes for inspection */
        for (uint64_t k = i + 1; k < n_particles; ++k) { c += norm[k]; cum[k] = c; }
    return clamped;
}

/* GridMapSlam::update, slam.rs:46-75 with externalised draws:
 * z = 2*N standard normals (per particle: centre draw, heading draw), u01 = resample uniform */
int so_update(struct so_slam* s, const double* angle, const double* dist, const uint8_t* valid, uint64_t nb,
              float dl, float dr, float wheel, const double* z, double u01) {
    double od[4];
    so_odometry_new(dl, dr, wheel, od);
    s->trace_count = 0;
    const int64_t n = (int64_t)s->n;
    /* ParticleFilter::update, particle.rs:31-35 (sequential in the reference; the optional
     * thread team only exists for the bench's all-cores baseline -- particles are independent) */
#pragma omp parallel for schedule(dynamic, 1) num_threads(s->threads) if (s->threads > 1)
    for (int64_t p = 0; p < n; ++p) particle_step(s, (uint64_t)p, angle, dist, valid, nb, od, z);

    if (s->carry)   /* adaptive resampling: weights accumulate over the steps that did not resample */
        for (int64_t p = 0; p < n; ++p) s->raw_weight[p] = s->carry[p] * s->raw_weight[p];
    memcpy(s->own_raw, s->raw_weight, s->n * sizeof(double));
    if (s->weight_override) {   /* resample on the weights the device computed (see so_set_weight_override) */
        memcpy(s->raw_weight, s->weight_override, s->n * sizeof(double));
        s->weight_override = NULL;
    }
    /* normalize_weights, argmax, resample indices: particle.rs:40-56, 78-101 */
    s->clamped = so_resample_fold(s->raw_weight, s->n, u01, s->weight, NULL, s->last_idx, &s->max_particle);
    s->resampled = 1;
    if (s->adaptive_tau > 0.0) {
        /* NOT in the reference (which resamples after every scan, slam.rs:74): resample only when the effective
         * number of particles (particle.rs:59-65) drops below tau * N; otherwise every particle stays where it
         * is and carries its normalised weight into the next update. */
        double sq = 0.0;
        for (int64_t p = 0; p < n; ++p) sq += s->weight[p] * s->weight[p];
        if (!(1.0 / sq < s->adaptive_tau * (double)s->n)) {
            if (!s->carry) s->carry = (double*)malloc(s->n * sizeof(double));
            memcpy(s->carry, s->weight, s->n * sizeof(double));
            for (int64_t p = 0; p < n; ++p) s->last_idx[p] = (uint64_t)p;
            s->clamped = 0;
            s->resampled = 0;
            return 0;
        }
        free(s->carry);
        s->carry = NULL;
    }
    /* new generation: clone(old[i]) for every slot (deep copy of Pose + Map) */
    struct so_map* new_map = (struct so_map*)calloc(s->n, sizeof(struct so_map));
    so_pose* new_pose = (so_pose*)malloc(s->n * sizeof(so_pose));
    if (!new_map || !new_pose) { free(new_map); free(new_pose); return -1; }
    int fail = 0;
#pragma omp parallel for schedule(static) num_threads(s->threads) if (s->threads > 1)
    for (int64_t mm = 0; mm < n; ++mm) {
        if (map_alloc(&new_map[mm], s)) { fail = 1; continue; }
        map_copy(&new_map[mm], &s->map[s->last_idx[mm]]);
        new_pose[mm] = s->pose[s->last_idx[mm]];
    }
    if (fail) return -1;
    for (int64_t p = 0; p < n; ++p) map_free(&s->map[p]);
    free(s->map); free(s->pose);
    s->map = new_map; s->pose = new_pose;
    return s->clamped ? 1 : 0;
}

/* ---------------------------------------------------------------- accessors */
uint64_t so_n(const struct so_slam* s) { return s->n; }
uint64_t so_grid_w(const struct so_slam* s) { return s->gw; }
uint64_t so_grid_h(const struct so_slam* s) { return s->gh; }
uint64_t so_max_particle(const struct so_slam* s) { return s->max_particle; }
void so_get_poses(const struct so_slam* s, float* out_xyt) {
    for (uint64_t i = 0; i < s->n; ++i) { out_xyt[3 * i] = s->pose[i].x; out_xyt[3 * i + 1] = s->pose[i].y; out_xyt[3 * i + 2] = s->pose[i].theta; }
}
void so_set_poses(struct so_slam* s, const float* xyt) {
    for (uint64_t i = 0; i < s->n; ++i) { s->pose[i].x = xyt[3 * i]; s->pose[i].y = xyt[3 * i + 1]; s->pose[i].theta = xyt[3 * i + 2]; }
}
void so_get_weights(const struct so_slam* s, double* norm, double* raw) {
    if (norm) memcpy(norm, s->weight, s->n * sizeof(double));
    if (raw) memcpy(raw, s->raw_weight, s->n * sizeof(double));
}
/* ParticleFilter::number_of_effective_particles, slamrs/slam/src/grid/particle.rs:59-65, over the
 * normalised weights of the last update (before resampling resets them to 1/N) */
double so_number_of_effective_particles(const struct so_slam* s) {
    double sum = 0.0;
    for (uint64_t i = 0; i < s->n; ++i) sum += s->weight[i] * s->weight[i];
    return 1.0 / sum;
}
void so_get_indices(const struct so_slam* s, uint64_t* idx) { memcpy(idx, s->last_idx, s->n * sizeof(uint64_t)); }
void so_get_odds(const struct so_slam* s, uint64_t particle, double* out) {
    const struct so_map* m = &s->map[particle];
    if (m->tiles) {
        for (uint64_t row = 0; row < s->gh; ++row)
            for (uint64_t col = 0; col < s->gw; ++col) out[cell_index(m, col, row)] = map_read_odds(m, col, row);
        return;
    }
    memcpy(out, m->odds, (size_t)(s->gw * s->gh) * sizeof(double));
}
/* Resample the NEXT update on these raw weights instead of the oracle's own (which stay readable through
 * so_get_own_raw). The device's exp/log differ from glibc's in the last bit, so its raw weights agree
 * with the oracle's to ~1e-12 relative, not bit for bit; with tens of thousands of particles a resampling
 * threshold lands that close to a prefix sum once in a few hundred steps and the two populations would
 * part ways. Lockstep tests at those sizes check the weights against the tolerance, then let both sides
 * resample on identical numbers -- where the index vector must match bit for bit. */
void so_set_weight_override(struct so_slam* s, const double* raw) { s->weight_override = raw; }
void so_get_own_raw(const struct so_slam* s, double* out) { memcpy(out, s->own_raw, s->n * sizeof(double)); }
int so_get_counts(const struct so_slam* s, uint64_t particle, uint16_t* n_free, uint16_t* n_occ) {
    if (s->sparse) {
        const struct so_map* m = &s->map[particle];
        for (uint64_t row = 0; row < s->gh; ++row)
            for (uint64_t col = 0; col < s->gw; ++col) {
                const struct so_tile* t = m->tiles[tile_of(m, col, row)];
                n_free[cell_index(m, col, row)] = t ? t->n_free[in_tile(col, row)] : 0;
                n_occ[cell_index(m, col, row)] = t ? t->n_occ[in_tile(col, row)] : 0;
            }
        return 0;
    }
    if (!s->track_counts) return -1;
    size_t cells = (size_t)(s->gw * s->gh);
    memcpy(n_free, s->map[particle].n_free, cells * 2);
    memcpy(n_occ, s->map[particle].n_occ, cells * 2);
    return 0;
}
/* GridMapSlam::estimated_pose, slam.rs:77-81 (indexes the NEW generation with the stale argmax) */
so_pose so_estimated_pose(const struct so_slam* s) { return s->pose[s->max_particle]; }
/* GridMapSlam::estimated_likelihood, slam.rs:83-88 */
void so_estimated_likelihood(const struct so_slam* s, double* out) {
    const struct so_map* m = &s->map[s->max_particle];
    size_t cells = (size_t)(s->gw * s->gh);
    if (m->tiles) {
        for (uint64_t row = 0; row < s->gh; ++row)
            for (uint64_t col = 0; col < s->gw; ++col)
                out[cell_index(m, col, row)] = so_log_odds_probability(map_read_odds(m, col, row));
        return;
    }
    for (size_t c = 0; c < cells; ++c) out[c] = so_log_odds_probability(m->odds[c]);
}
int64_t so_get_trace(const struct so_slam* s, int32_t* out, int64_t cap) {
    int64_t k = s->trace_count < s->trace_cap ? s->trace_count : s->trace_cap;
    if (k > cap) k = cap;
    if (out && k > 0) memcpy(out, s->trace, (size_t)k * 3 * sizeof(int32_t));
    return s->trace_count;
}
int so_clamped(const struct so_slam* s) { return s->clamped; }

/* ---------------------------------------------------------------- scan generator
 * Restates the simulator's lidar (slamrs/simulator/src/sim.rs:134-159) against line segments
 * (slamrs/simulator/src/scene/ray.rs:55-83, 164-172) and its motion model (sim.rs:214-220). */
static int seg_intersect(const float* seg, float ox, float oy, float dx, float dy, float* u_out) {
    float x1 = seg[0], y1 = seg[1], x2 = seg[2], y2 = seg[3];
    float x3 = ox, y3 = oy, x4 = ox + dx, y4 = oy + dy;
    float denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (denom == 0.0f) return 0;
    float t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / denom;
    float u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / denom;
    if (t >= 0.0f && t <= 1.0f && u > 0.0f) { *u_out = u; return 1; }
    return 0;
}

/* returns number of measurements produced (beams whose ray hits nothing are dropped, sim.rs:138) */
uint64_t so_sim_scan(const float* segments, uint64_t n_seg, float px, float py, float ptheta, uint64_t n_beams,
                     float scanner_range, double* angle, double* dist, uint8_t* valid) {
    uint64_t k = 0;
    for (uint64_t b = 0; b < n_beams; ++b) {
        /* (angle as f32).to_radians(): value * (PI_f32 / 180) ; generalised to 360/n_beams degree steps */
        float deg = (float)b * (360.0f / (float)n_beams);
        float a = deg * (3.14159265358979323846264338327950288f / 180.0f);
        float dir = a + ptheta;
        float dx = cosf(dir), dy = sinf(dir);
        int have = 0;
        float best = 0.0f;
        for (uint64_t sidx = 0; sidx < n_seg; ++sidx) {
            float u;
            if (seg_intersect(&segments[4 * sidx], px, py, dx, dy, &u)) {
                /* min_by(partial_cmp().unwrap_or(Less)): keeps the earlier element on ties */
                if (!have || u < best) { best = u; have = 1; }
            }
        }
        if (!have) continue;
        angle[k] = (double)a;
        if (best < scanner_range) { dist[k] = (double)best; valid[k] = 1; }
        else { dist[k] = (double)scanner_range; valid[k] = 0; }
        k++;
    }
    return k;
}

/* Simulator::motion_model, sim.rs:214-220 */
void so_sim_motion(float* px, float* py, float* ptheta, float sl, float sr, float wheel_base) {
    float sbar = (sr + sl) / 2.0f;
    *ptheta += (sr - sl) / wheel_base;
    *px += sbar * cosf(*ptheta);
    *py += sbar * sinf(*ptheta);
}

/* the platform libm's f32 sin/cos -- what Rust's f32::sin / f32::cos call on x86-64 Linux */
void so_libm_sincosf(const float* x, uint64_t n, float* s, float* c) {
    for (uint64_t i = 0; i < n; ++i) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); }
}
