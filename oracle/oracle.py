"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (slamrs_b200/) never does. See oracle/slam_oracle.c for
the parity status ("parity unpinned" for the grid module) and the reference file:line map.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -O3 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("slam_oracle.c", "shared_stream.c", "slam_oracle.h", "shared_stream.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Pose(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("theta", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    d, f, u64, i64, vp = C.c_double, C.c_float, C.c_uint64, C.c_int64, C.c_void_p
    L.so_prob_log_odds.restype = d; L.so_prob_log_odds.argtypes = [d]
    L.so_log_odds_probability.restype = d; L.so_log_odds_probability.argtypes = [d]
    L.so_angle_diff.restype = d; L.so_angle_diff.argtypes = [d, d]
    L.so_odometry_new.restype = None; L.so_odometry_new.argtypes = [f, f, f, vp]
    L.so_odometry_sample.restype = _Pose; L.so_odometry_sample.argtypes = [vp, _Pose, d, d]
    L.so_odometry_log_prob.restype = d; L.so_odometry_log_prob.argtypes = [vp, _Pose, _Pose]
    L.so_ray_cells.restype = i64; L.so_ray_cells.argtypes = [f, f, f, f, u64, u64, u64, vp, i64]
    L.so_inverse_sensor_model.restype = C.c_int; L.so_inverse_sensor_model.argtypes = [f, f, C.c_int, f]
    L.so_grid_cells.restype = u64; L.so_grid_cells.argtypes = [f, f]
    L.so_create.restype = vp; L.so_create.argtypes = [f, f, f, f, f, u64, C.c_int]
    L.so_create_ex.restype = vp; L.so_create_ex.argtypes = [f, f, f, f, f, u64, C.c_int, C.c_int]
    L.so_set_weight_override.restype = None; L.so_set_weight_override.argtypes = [vp, vp]
    L.so_get_own_raw.restype = None; L.so_get_own_raw.argtypes = [vp, vp]
    L.so_destroy.restype = None; L.so_destroy.argtypes = [vp]
    L.so_set_adaptive_resampling.restype = None; L.so_set_adaptive_resampling.argtypes = [vp, d]
    L.so_resampled.restype = C.c_int; L.so_resampled.argtypes = [vp]
    L.ss_fill_uniform_poses.restype = None; L.ss_fill_uniform_poses.argtypes = [u64, u64, u64, d, d, d, d, vp]
    L.so_set_threads.restype = None; L.so_set_threads.argtypes = [vp, C.c_int]
    L.so_set_dead_likelihood.restype = None; L.so_set_dead_likelihood.argtypes = [vp, C.c_int]
    L.so_set_trace.restype = None; L.so_set_trace.argtypes = [vp, i64, i64]
    L.so_resample_fold.restype = C.c_int; L.so_resample_fold.argtypes = [vp, u64, d, vp, vp, vp, vp]
    L.so_update.restype = C.c_int; L.so_update.argtypes = [vp, vp, vp, vp, u64, f, f, f, vp, d]
    for name in ("so_n", "so_grid_w", "so_grid_h", "so_max_particle"):
        getattr(L, name).restype = u64; getattr(L, name).argtypes = [vp]
    L.so_get_poses.restype = None; L.so_get_poses.argtypes = [vp, vp]
    L.so_set_poses.restype = None; L.so_set_poses.argtypes = [vp, vp]
    L.so_get_weights.restype = None; L.so_get_weights.argtypes = [vp, vp, vp]
    L.so_get_indices.restype = None; L.so_get_indices.argtypes = [vp, vp]
    L.so_number_of_effective_particles.restype = C.c_double; L.so_number_of_effective_particles.argtypes = [vp]
    L.so_get_odds.restype = None; L.so_get_odds.argtypes = [vp, u64, vp]
    L.so_get_counts.restype = C.c_int; L.so_get_counts.argtypes = [vp, u64, vp, vp]
    L.so_estimated_pose.restype = _Pose; L.so_estimated_pose.argtypes = [vp]
    L.so_estimated_likelihood.restype = None; L.so_estimated_likelihood.argtypes = [vp, vp]
    L.so_get_trace.restype = i64; L.so_get_trace.argtypes = [vp, vp, i64]
    L.so_clamped.restype = C.c_int; L.so_clamped.argtypes = [vp]
    L.so_sim_scan.restype = u64; L.so_sim_scan.argtypes = [vp, u64, f, f, f, u64, f, vp, vp, vp]
    L.so_sim_motion.restype = None; L.so_sim_motion.argtypes = [vp, vp, vp, f, f, f]
    L.so_libm_sincosf.restype = None; L.so_libm_sincosf.argtypes = [vp, u64, vp, vp]
    L.ss_philox4x32_10.restype = None; L.ss_philox4x32_10.argtypes = [vp, vp, vp]
    L.ss_dlog.restype = d; L.ss_dlog.argtypes = [d]
    L.ss_dsincos2pi.restype = None; L.ss_dsincos2pi.argtypes = [d, vp, vp]
    L.ss_fill_motion_normals.restype = None; L.ss_fill_motion_normals.argtypes = [u64, u64, u64, u64, vp]
    L.ss_resample_uniform.restype = d; L.ss_resample_uniform.argtypes = [u64, u64]
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- scalar pieces
def prob_log_odds(p): return lib().so_prob_log_odds(p)
def log_odds_probability(l): return lib().so_log_odds_probability(l)
def angle_diff(a, b): return lib().so_angle_diff(a, b)
def grid_cells(extent, res): return int(lib().so_grid_cells(extent, res))
def inverse_sensor_model(d, md, hit, tol=2.0): return int(lib().so_inverse_sensor_model(d, md, int(hit), tol))


def odometry_new(dl, dr, wheel) -> np.ndarray:
    out = np.zeros(4, np.float64)
    lib().so_odometry_new(dl, dr, wheel, _p(out))
    return out


def odometry_sample(od, pose, z1, z2):
    od = np.ascontiguousarray(od, np.float64)
    r = lib().so_odometry_sample(_p(od), _Pose(*[float(v) for v in pose]), z1, z2)
    return np.array([r.x, r.y, r.theta], np.float32)


def odometry_log_prob(od, a, b):
    od = np.ascontiguousarray(od, np.float64)
    return lib().so_odometry_log_prob(_p(od), _Pose(*[float(v) for v in a]), _Pose(*[float(v) for v in b]))


def ray_cells(x0, y0, x1, y1, w, h, extra=2) -> np.ndarray:
    cap = int(w + h + 16 + extra)
    out = np.zeros((cap, 2), np.int32)
    n = lib().so_ray_cells(x0, y0, x1, y1, w, h, extra, _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def libm_sincosf(x):
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    s = np.zeros_like(x); c = np.zeros_like(x)
    lib().so_libm_sincosf(_p(x), x.size, _p(s), _p(c))
    return s, c


# ---------------------------------------------------------------- shared stream
def philox(ctr, key) -> np.ndarray:
    c = np.ascontiguousarray(ctr, np.uint32); k = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().ss_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def dlog(x): return lib().ss_dlog(x)


def dsincos2pi(u):
    s = C.c_double(); c = C.c_double()
    lib().ss_dsincos2pi(u, C.byref(s), C.byref(c))
    return s.value, c.value


def motion_normals(seed: int, step: int, first: int, count: int) -> np.ndarray:
    z = np.zeros(2 * count, np.float64)
    lib().ss_fill_motion_normals(seed, step, first, count, _p(z))
    return z


def uniform_poses(seed: int, first: int, count: int, box) -> np.ndarray:
    out = np.zeros((count, 3), np.float32)
    lib().ss_fill_uniform_poses(seed, first, count, float(box[0]), float(box[1]), float(box[2]), float(box[3]), _p(out))
    return out


def resample_uniform(seed: int, step: int) -> float:
    return lib().ss_resample_uniform(seed, step)


# ---------------------------------------------------------------- simulator restatement
def resample_fold(raw_weights, u01: float):
    """normalize_weights + argmax + resample indices (particle.rs:40-56, 78-101) on given raw weights:
    dict(norm, cum, idx, max_particle, clamped)."""
    w = np.ascontiguousarray(raw_weights, np.float64).reshape(-1)
    n = w.size
    norm = np.zeros(n, np.float64); cum = np.zeros(n, np.float64); idx = np.zeros(n, np.uint64)
    mp = C.c_uint64(0)
    clamped = lib().so_resample_fold(_p(w), n, float(u01), _p(norm), _p(cum), _p(idx), C.byref(mp))
    return dict(norm=norm, cum=cum, idx=idx, max_particle=int(mp.value), clamped=int(clamped))


def sim_scan(segments, pose, n_beams, scanner_range):
    seg = np.ascontiguousarray(segments, np.float32).reshape(-1, 4)
    angle = np.zeros(n_beams, np.float64); dist = np.zeros(n_beams, np.float64); valid = np.zeros(n_beams, np.uint8)
    k = lib().so_sim_scan(_p(seg), seg.shape[0], pose[0], pose[1], pose[2], n_beams, scanner_range,
                          _p(angle), _p(dist), _p(valid))
    return angle[:k].copy(), dist[:k].copy(), valid[:k].copy()


def sim_motion(pose, sl, sr, wheel_base):
    x = C.c_float(pose[0]); y = C.c_float(pose[1]); t = C.c_float(pose[2])
    lib().so_sim_motion(C.byref(x), C.byref(y), C.byref(t), sl, sr, wheel_base)
    return (x.value, y.value, t.value)


# ---------------------------------------------------------------- the filter
class OracleSlam:
    """GridMapSlam (slam.rs:13-97) restated on the CPU with externalised random draws."""

    def __init__(self, position, width, height, resolution, n_particles, track_counts=True, sparse=False):
        """sparse: tile storage with copy-on-write clones (same arithmetic; for populations whose dense f64
        grids would not fit in host memory)."""
        self._h = lib().so_create_ex(position[0], position[1], width, height, resolution, n_particles, int(track_counts),
                                     int(sparse))
        self._override = None
        if not self._h:
            raise MemoryError("oracle allocation failed")
        self.n = int(lib().so_n(self._h))
        self.gw = int(lib().so_grid_w(self._h))
        self.gh = int(lib().so_grid_h(self._h))

    def close(self):
        if self._h:
            lib().so_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_threads(self, t): lib().so_set_threads(self._h, t)
    def set_adaptive_resampling(self, tau): lib().so_set_adaptive_resampling(self._h, float(tau))
    def resampled(self): return bool(lib().so_resampled(self._h))
    def set_dead_likelihood(self, on): lib().so_set_dead_likelihood(self._h, int(on))
    def set_trace(self, particle, cap): lib().so_set_trace(self._h, particle, cap)

    def set_weight_override(self, raw):
        """The next update resamples on these raw weights (see so_set_weight_override in slam_oracle.c)."""
        self._override = np.ascontiguousarray(raw, np.float64).copy()
        assert self._override.size == self.n
        lib().so_set_weight_override(self._h, _p(self._override))

    def own_raw_weights(self):
        out = np.zeros(self.n, np.float64); lib().so_get_own_raw(self._h, _p(out)); return out

    def update(self, angle, dist, valid, dl, dr, wheel, z, u01) -> int:
        angle = np.ascontiguousarray(angle, np.float64); dist = np.ascontiguousarray(dist, np.float64)
        valid = np.ascontiguousarray(valid, np.uint8); z = np.ascontiguousarray(z, np.float64)
        assert z.size == 2 * self.n and angle.size == dist.size == valid.size
        return lib().so_update(self._h, _p(angle), _p(dist), _p(valid), angle.size, dl, dr, wheel, _p(z), u01)

    @property
    def max_particle(self): return int(lib().so_max_particle(self._h))

    def poses(self):
        out = np.zeros((self.n, 3), np.float32); lib().so_get_poses(self._h, _p(out)); return out

    def set_poses(self, xyt):
        a = np.ascontiguousarray(xyt, np.float32).reshape(self.n, 3); lib().so_set_poses(self._h, _p(a))

    def weights(self):
        w = np.zeros(self.n, np.float64); r = np.zeros(self.n, np.float64)
        lib().so_get_weights(self._h, _p(w), _p(r)); return w, r

    def number_of_effective_particles(self):
        return float(lib().so_number_of_effective_particles(self._h))

    def indices(self):
        i = np.zeros(self.n, np.uint64); lib().so_get_indices(self._h, _p(i)); return i

    def odds(self, particle):
        o = np.zeros(self.gw * self.gh, np.float64); lib().so_get_odds(self._h, particle, _p(o)); return o

    def counts(self, particle):
        a = np.zeros(self.gw * self.gh, np.uint16); b = np.zeros(self.gw * self.gh, np.uint16)
        rc = lib().so_get_counts(self._h, particle, _p(a), _p(b)); assert rc == 0
        return a, b

    def estimated_pose(self):
        r = lib().so_estimated_pose(self._h); return np.array([r.x, r.y, r.theta], np.float32)

    def estimated_likelihood(self):
        o = np.zeros(self.gw * self.gh, np.float64); lib().so_estimated_likelihood(self._h, _p(o)); return o

    def trace(self):
        n = lib().so_get_trace(self._h, None, 0)
        out = np.zeros((max(n, 1), 3), np.int32)
        lib().so_get_trace(self._h, _p(out), n)
        return out[:n]

    @property
    def clamped(self): return bool(lib().so_clamped(self._h))
