"""TEST INFRASTRUCTURE ONLY -- a second, independent restatement of the reference step.

Written from the prose specification in SURVEY.md Appendix A (A.1-A.8), NOT from
oracle/slam_oracle.c, in pure Python with numpy scalar types, so that a misreading of the Rust
source would have to be made twice, in two different shapes, to go unnoticed. Slow: small cases
only (tests/test_oracle_cross_check.py). f32 sin/cos/sqrt come from the platform libm through
ctypes (what Rust's f32 methods call on x86-64 Linux), f64 exp/log from Python's math module
(the same libm).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.sinf.restype = ctypes.c_float; _libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float; _libm.cosf.argtypes = [ctypes.c_float]

F = np.float32
PI = math.pi


def sinf(x): return F(_libm.sinf(float(x)))
def cosf(x): return F(_libm.cosf(float(x)))


def rust_as_usize(v) -> int:
    v = float(v)
    if v != v or v <= 0.0:
        return 0
    if v >= 2.0 ** 64:
        return 2 ** 64 - 1
    return int(v)


def rust_as_isize(v) -> int:
    v = float(v)
    if v != v:
        return 0
    if v >= 2.0 ** 63:
        return 2 ** 63 - 1
    if v <= -(2.0 ** 63):
        return -(2 ** 63)
    return int(v)


L_FREE = math.log(0.30 / (1.0 - 0.30))
L_OCC = math.log(0.9 / (1.0 - 0.9))
L_PRIOR = math.log(0.5 / (1.0 - 0.5))


def probability(odds: float) -> float:          # A.4
    try:
        e = math.exp(odds)
    except OverflowError:
        e = math.inf
    return 1.0 - 1.0 / (1.0 + e)


def angle_diff(alpha: float, beta: float) -> float:   # A.5
    diff = math.fmod(beta - alpha + PI, PI * 2.0) - PI
    return diff + 2.0 * PI if diff < -PI else diff


def normal_pdf(x, mean, sd):
    d = (x - mean) / sd
    # statrs consts::SQRT_2PI is a 50-digit decimal literal; its nearest f64 is 0x1.40d931ff62706p+1
    # (NOT sqrt(2*pi) evaluated in f64, which is one ulp lower)
    return math.exp(-0.5 * d * d) / (float("2.5066282746310005024157652848110452530069867406099") * sd)


def ray_cells(x0, y0, x1, y1, w, h, extra=2):
    """A.7 -- list of (X, Y) in emission order."""
    x0, y0, x1, y1 = F(x0), F(y0), F(x1), F(y1)
    with np.errstate(all="ignore"):
        dx, dy = np.abs(F(x1 - x0)), np.abs(F(y1 - y0))
        X, Y = rust_as_isize(np.floor(x0)), rust_as_isize(np.floor(y0))
        n = 1 + extra
        if dx == 0:
            xi, err = 0, F(np.inf)
        elif x1 > x0:
            xi = 1
            n += rust_as_isize(F(np.floor(x1) - F(X)))
            err = F(F(F(np.floor(x0) + F(1.0)) - x0) * dy)
        else:
            xi = -1
            n += X - rust_as_isize(np.floor(x1))
            err = F(F(x0 - np.floor(x0)) * dy)
        if dy == 0:
            yi = 0
            err = F(err - F(np.inf))
        elif y1 > y0:
            yi = 1
            n += rust_as_isize(np.floor(y1)) - Y
            err = F(err - F(F(F(np.floor(y0) + F(1.0)) - y0) * dx))
        else:
            yi = -1
            n += Y - rust_as_isize(np.floor(y1))
            err = F(err - F(F(y0 - np.floor(y0)) * dx))
        out = []
        n &= (1 << 64) - 1
        while n > 0 and 0 <= X < w and 0 <= Y < h:
            out.append((X, Y))
            if err > 0:
                Y += yi
                err = F(err - dx)
            else:
                X += xi
                err = F(err + dy)
            n -= 1
        return out


def sensor_increment(dd, md, hit):   # A.6
    if not hit:
        return L_FREE if dd < md else L_PRIOR
    if dd < F(md - F(1.0)):
        return L_FREE
    if dd > F(md + F(1.0)):
        return L_PRIOR
    return L_OCC


class Slam:
    def __init__(self, position, width, height, resolution, n):
        self.pos = (F(position[0]), F(position[1]))
        self.res = F(resolution)
        self.w = rust_as_usize(np.ceil(F(F(width) / self.res)))      # A.1
        self.h = rust_as_usize(np.ceil(F(F(height) / self.res)))
        self.n = n
        self.pose = [(F(0), F(0), F(0)) for _ in range(n)]
        self.grid = [np.zeros(self.w * self.h, np.float64) for _ in range(n)]
        self.max_particle = 0
        self.weights = [1.0 / n] * n
        self.idx = list(range(n))

    def to_grid(self, wx, wy):
        return F(F(wx - self.pos[0]) / self.res), F(F(wy - self.pos[1]) / self.res)

    def endpoint(self, pose, angle, dist):
        a = F(pose[2] + F(angle))
        d = F(dist)
        return F(pose[0] + F(cosf(a) * d)), F(pose[1] + F(sinf(a) * d))

    def update(self, angle, dist, valid, dl, dr, wheel, z, u01):
        dl, dr, wheel = F(dl), F(dr), F(wheel)
        mu_c = float(F(F(dl + dr) / F(2.0)))                         # A.2
        mu_t = float(F(F(dr - dl) / wheel))
        sd_c = (0.01 + abs(mu_c) * 0.05) / 2.0
        sd_t = 5.0 * (PI / 180.0) + 0.1 * abs(mu_t)
        raw = []
        for p in range(self.n):
            x, y, th = self.pose[p]
            d = F(mu_c + sd_c * z[2 * p])                             # A.3
            nth = F(th + F(mu_t + sd_t * z[2 * p + 1]))
            nx = F(x + F(cosf(nth) * d))
            ny = F(y + F(sinf(nth) * d))
            g = self.grid[p]
            L = math.log(1.0)                                         # A.4
            for a, r, v in zip(angle, dist, valid):
                if not v:
                    continue
                ex, ey = self.endpoint((nx, ny, nth), a, r)
                gx, gy = self.to_grid(ex, ey)
                if gx < 0 or gy < 0 or rust_as_usize(gx) >= self.w or rust_as_usize(gy) >= self.h:
                    continue
                pr = probability(g[rust_as_usize(gy) * self.h + rust_as_usize(gx)])
                L += math.log(1.0 / 1.0) if pr == 0.5 else math.log(0.9 * pr + (1.0 - 0.9) * 1.0 / 1.0)
            cd = F(np.sqrt(F(F(F(x - nx) * F(x - nx)) + F(F(y - ny) * F(y - ny)))))   # A.5
            M = math.log(normal_pdf(float(cd), mu_c, sd_c)) + math.log(normal_pdf(angle_diff(float(th), float(nth)), mu_t, sd_t))
            sx, sy = self.to_grid(nx, ny)                             # A.6
            for a, r, v in zip(angle, dist, valid):
                ex, ey = self.endpoint((nx, ny, nth), a, r)
                gx, gy = self.to_grid(ex, ey)
                md = F(F(r) / self.res)
                for (X, Y) in ray_cells(sx, sy, gx, gy, self.w, self.h, 2):
                    ddx = F(sx - F(F(X) + F(0.5)))
                    ddy = F(sy - F(F(Y) + F(0.5)))
                    dd = F(np.sqrt(F(F(F(0) + F(ddx * ddx)) + F(ddy * ddy))))
                    g[Y * self.h + X] += sensor_increment(dd, md, bool(v))
            self.pose[p] = (nx, ny, nth)
            try:
                raw.append(math.exp(L + M))
            except OverflowError:
                raw.append(math.inf)
        total = 0.0                                                   # A.8
        for w in raw:
            total += w
        wn = [w / total for w in raw]
        best = 0
        for i in range(1, self.n):
            if _total_key(wn[i]) >= _total_key(wn[best]):
                best = i
        self.max_particle = best
        self.raw = raw
        self.weights = wn
        N = float(self.n)
        r = u01 * 1.0 / N
        c = wn[0]
        i = 0
        idx = []
        for m in range(1, self.n + 1):
            u = r + (float(m) - 1.0) * 1.0 / N
            while u > c and i + 1 < self.n:
                i += 1
                c += wn[i]
            idx.append(i)
        self.idx = idx
        self.grid = [self.grid[i].copy() for i in idx]
        self.pose = [self.pose[i] for i in idx]


def _total_key(v: float) -> int:
    b = int(np.float64(v).view(np.int64))
    return b ^ (((b >> 63) & ((1 << 64) - 1)) >> 1) if b < 0 else b
