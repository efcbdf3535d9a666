"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bit-exact: sin/cos, shared stream, ray-cast cell sequences, poses, hit counters, resample
indices, argmax. Within tolerance (1e-9 relative, contract is 1e-5): weights, log-odds, map.
"""
import numpy as np
import pytest

from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig, Observation, Odometry
from slamrs_b200 import _lib
from slamrs_b200 import slam as S

from common import SEED, compare_step, lockstep, make_scans, oracle_slam, oracle_step

pytestmark = pytest.mark.gpu


def test_device_sincos_is_glibc_bit_exact(oracle):
    rng = np.random.default_rng(1)
    xs = np.concatenate([
        rng.uniform(-10, 10, 400_000), rng.uniform(-130, 130, 300_000), rng.uniform(-1e-3, 1e-3, 50_000),
        rng.uniform(-1e6, 1e6, 100_000), rng.standard_normal(100_000) * 1e20,
        np.array([0.0, -0.0, np.pi / 4, -np.pi / 4, 0.78539816, 0.7853982, 119.99999, 120.0, 120.00001, 3.4e38, 1e-45]),
    ]).astype(np.float32)
    s_ref, c_ref = oracle.libm_sincosf(xs)
    s, c = S.debug_sincos(xs)
    assert np.array_equal(s.view(np.uint32), s_ref.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), c_ref.view(np.uint32))


def test_device_shared_stream_is_bit_exact(oracle):
    for step, first, count in [(0, 0, 4096), (7, 12345, 1000), (2**33 + 5, 65000, 600)]:
        z_ref = oracle.motion_normals(SEED, step, first, count)
        u_ref = oracle.resample_uniform(SEED, step)
        z, u = S.debug_stream(SEED, step, first, count)
        assert np.array_equal(z.view(np.uint64), z_ref.view(np.uint64))
        assert u == u_ref


def _ray_cases(rng, w, h, n):
    x0 = rng.uniform(-2, w + 2, n); y0 = rng.uniform(-2, h + 2, n)
    ang = rng.uniform(0, 2 * np.pi, n); ln = rng.uniform(0, 1.5 * w, n)
    x1 = x0 + np.cos(ang) * ln; y1 = y0 + np.sin(ang) * ln
    # axis-aligned, zero-length, integer coordinates, starts on cell borders
    k = n // 8
    x1[:k] = x0[:k]; y1[k:2 * k] = y0[k:2 * k]
    x1[2 * k:3 * k] = x0[2 * k:3 * k]; y1[2 * k:3 * k] = y0[2 * k:3 * k]
    x0[3 * k:4 * k] = np.floor(x0[3 * k:4 * k]); y0[4 * k:5 * k] = np.floor(y0[4 * k:5 * k])
    return [v.astype(np.float32) for v in (x0, y0, x1, y1)]


@pytest.mark.parametrize("w,h", [(200, 200), (64, 64), (1024, 1024)])
def test_raycast_cells_bit_exact(oracle, w, h):
    rng = np.random.default_rng(w)
    x0, y0, x1, y1 = _ray_cases(rng, w, h, 4000)
    cells, counts = S.debug_raycast(x0, y0, x1, y1, w, h, extra=2)
    for i in range(x0.size):
        ref = oracle.ray_cells(x0[i], y0[i], x1[i], y1[i], w, h, 2)
        assert counts[i] == len(ref), (i, x0[i], y0[i], x1[i], y1[i])
        assert np.array_equal(cells[i, :counts[i]], ref), (i, x0[i], y0[i], x1[i], y1[i])


BOTH_RAY_KERNELS = pytest.mark.parametrize("flags", [0, _lib.FLAG_GENERIC_RAY_KERNEL], ids=["packed-window", "generic-window"])


@BOTH_RAY_KERNELS
def test_step_parity_default_preset(oracle, flags):
    """configs[0]: the shipped preset (200x200 @ 2 cm, 1 m range), 30 particles, 8 scans."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=30)
    errs = lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 8), flags=flags)
    print(errs[-1])


def test_step_parity_caller_supplied_draws(oracle):
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=10)
    lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 4), rng_mode=_lib.RNG_CALLER)


@BOTH_RAY_KERNELS
def test_step_parity_long_range(oracle, flags):
    """configs[1] geometry: 6 m range at 5 cm cells (disc window of radius 124 cells)."""
    cfg = GridMapSlamConfig(position=(-12.8, -12.8), width=25.6, height=25.6, resolution=0.05, n_particles=24)
    scans = make_scans(5.0, 360, 6.0, 4)
    errs = lockstep(oracle, cfg, scans, particles=range(0, 24, 5), flags=flags)
    print(errs[-1])


@BOTH_RAY_KERNELS
def test_step_parity_range_beyond_window_spills(oracle, flags):
    """Up to 7 m rays at 2.5 cm cells = 280-cell rays: longer than any shared-memory window, so the
    exact global-memory path handles the tails; counters must still be exact."""
    cfg = GridMapSlamConfig(position=(-12.8, -12.8), width=25.6, height=25.6, resolution=0.025, n_particles=6)
    scans = make_scans(5.0, 360, 12.0, 3)
    with GridMapSlam(cfg, GpuPlacement(flags=flags)) as g:
        g.update(*scans[0])
        assert g.stats()["spilled_cells"] > 0
    lockstep(oracle, cfg, scans, particles=[0, 3, 5], flags=flags)


@BOTH_RAY_KERNELS
def test_many_hits_in_one_cell(oracle, flags):
    """Adversarial scan: every beam at the same angle and distance, so single cells collect
    hundreds of free and occupied hits in one scan (overflows the packed window's 5-bit field)."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=4)
    n = 360
    same = Observation(0, angle=np.full(n, 0.3), distance=np.full(n, 0.5), valid=np.ones(n, bool))
    near = Observation(0, angle=np.linspace(0, 2 * np.pi, n, endpoint=False), distance=np.full(n, 0.03), valid=np.ones(n, bool))
    odo = Odometry.new(0.01, 0.012, 0.1)
    lockstep(oracle, cfg, [(same, odo), (near, odo), (same, odo)], flags=flags)


COPY_MODES = pytest.mark.parametrize("copy_flags", [0, _lib.FLAG_EAGER_COPY, _lib.FLAG_FULL_GRID_COPY],
                                     ids=["deferred-extent-copy", "eager-extent-copy", "whole-grid-copy"])


@COPY_MODES
def test_step_parity_scattered_poses_mixed_extents(oracle, copy_flags):
    """Particles are re-scattered over the room before every scan, so the slots the resampler
    recycles hold grids with very different informed extents: an extent-limited copy must also
    clear whatever the destination's previous tenant had informed outside the source's extent."""
    cfg = GridMapSlamConfig(position=(-6.4, -6.4), width=12.8, height=12.8, resolution=0.05, n_particles=48)
    scans = make_scans(5.0, 360, 3.0, 6)
    rng = np.random.default_rng(7)

    def scatter(step, gpu, osl):
        xyt = np.column_stack([rng.uniform(-4.5, 4.5, 48), rng.uniform(-4.5, 4.5, 48),
                               rng.uniform(-np.pi, np.pi, 48)]).astype(np.float32)
        if step % 2 == 0:
            gpu.set_poses(xyt); osl.set_poses(xyt)

    errs = lockstep(oracle, cfg, scans, flags=copy_flags, pre_step=scatter)
    print(errs[-1])


@pytest.mark.parametrize("slot_cells", [0, 256], ids=["whole-grid-slots", "windowed-slots"])
def test_band_extents_cover_exactly_what_is_informed(oracle, slot_cells):
    """The resampler's view of every grid -- bounding box and per-band column ranges -- must contain every
    informed cell, stay inside the box, and be empty for every band outside the box's rows (the invariant
    the extent copy relies on). Scattered start poses, so slots change tenants with different extents."""
    cfg = GridMapSlamConfig(position=(-10.24, -10.24), width=20.48, height=20.48, resolution=0.04, n_particles=32)
    assert S.grid_cells(20.48, 0.04) == 512
    scans = make_scans(2.0, 360, 2.0, 5)
    rng = np.random.default_rng(21)
    init = np.column_stack([rng.uniform(-8.0, 8.0, 32), rng.uniform(-8.0, 8.0, 32), rng.uniform(-np.pi, np.pi, 32)]).astype(np.float32)
    with GridMapSlam(cfg, GpuPlacement(slot_cells=slot_cells)) as g:
        g.set_poses(init)
        ph = slot_cells or 512
        for obs, odo in scans:
            g.update(obs, odo)
            for p in range(0, 32, 3):
                cells = g.cells(p).reshape(512, 512)
                (x0, y0, x1, y1), shift, bands = g.extents(p)
                assert bands.shape[0] == ph // 8 and shift % 8 == 0
                ys, xs = np.nonzero(cells)
                assert ys.size and x0 <= xs.min() and xs.max() < x1 and y0 <= ys.min() and ys.max() < y1
                assert x0 % 8 == 0 and x1 % 8 == 0 and x1 - x0 <= ph and y1 - y0 <= ph
                in_box = np.zeros(ph // 8, bool)
                for y in range(y0, y1):
                    in_box[(y % ph) // 8] = True
                    row = np.nonzero(cells[y])[0]
                    b0, b1 = bands[(y % ph) // 8]
                    if row.size:
                        assert b0 <= row.min() and row.max() < b1, (p, y)
                    assert (b0 == 0 and b1 == 0) or (x0 <= b0 < b1 <= x1)
                assert np.all(bands[~in_box] == 0), "a band outside the box's rows is not empty"
                # the ranges are tight enough to matter: well below the box area
                assert (bands[:, 1] - bands[:, 0]).sum() * 8 <= (x1 - x0) * (((y1 + 7) // 8 - y0 // 8) * 8)


def test_extent_copy_moves_fewer_bytes_same_result():
    """Both copy modes give identical grids; the extent-limited one reports the bytes it moved."""
    cfg = GridMapSlamConfig(position=(-12.8, -12.8), width=25.6, height=25.6, resolution=0.05, n_particles=256)
    scans = make_scans(5.0, 360, 6.0, 4)
    out = {}
    for flags in (0, _lib.FLAG_EAGER_COPY, _lib.FLAG_FULL_GRID_COPY):
        with GridMapSlam(cfg, GpuPlacement(flags=flags)) as g:
            for obs, odo in scans:
                g.update(obs, odo)
            st = g.stats()
            hist = g.step_history(0, len(scans))
            assert hist[-1, 5] == st["copy_bytes"]
            out[flags] = (st, [g.cells(p).copy() for p in (0, 1, 100, 255)], g.estimated_likelihood().data.copy())
    (st_box, cells_box, map_box), (st_full, cells_full, map_full) = out[_lib.FLAG_EAGER_COPY], out[_lib.FLAG_FULL_GRID_COPY]
    st_def, cells_def, map_def = out[0]
    for a, b, c in zip(cells_box, cells_full, cells_def):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(map_box, map_full) and np.array_equal(map_box, map_def)
    assert st_box["grids_copied"] == st_full["grids_copied"] > 0
    assert 0 < st_box["copy_bytes"] < st_full["copy_bytes"] // 4
    # deferred copies: only the clones that survive the next resampling are ever copied
    assert 0 < st_def["grids_copied"] <= st_def["particles_integrated"] < st_box["grids_copied"]
    assert 0 < st_def["copy_bytes"] < st_box["copy_bytes"]
    # whole-grid mode: every copy writes a grid, every fan-out sub-run reads one
    assert st_full["copy_bytes"] >= st_full["grids_copied"] * st_full["bytes_per_grid"]


def test_set_cells_then_resample_keeps_extents_consistent(oracle):
    """set_cells installs an arbitrary image (extent recomputed on the host); later copies of that
    particle and into its slot must reproduce it exactly."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=8)
    scans = make_scans(1.0, 360, 1.0, 3)
    rng = np.random.default_rng(3)
    img = np.zeros((200, 200), np.uint32)
    img[5:190, 3:199] = rng.integers(0, 4, (185, 196)).astype(np.uint32) * 0x10001
    with GridMapSlam(cfg) as g:
        g.update(*scans[0])
        for p in range(8):
            g.set_cells(p, img if p % 2 == 0 else np.zeros_like(img))
        for p in range(8):
            assert np.array_equal(g.cells(p).reshape(200, 200), img if p % 2 == 0 else 0 * img)
        g.update(*scans[1])
        idx = g.resample_indices()
        g.update(*scans[2])
        # every particle descends from one of the two images plus the same two scans at its own poses;
        # all cells outside the scans' reach must still be exactly the installed image
        for p in range(8):
            c = g.cells(p).reshape(200, 200)
            far = np.ones((200, 200), bool); far[40:160, 40:160] = False
            assert np.array_equal(c[far], img[far]) or np.array_equal(c[far], 0 * img[far])


def test_step_parity_720_beams_odd_grid(oracle):
    """720 beams and a grid whose side is not a multiple of 4 (scalar write-back path)."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.02, height=4.02, resolution=0.02, n_particles=12)
    assert S.grid_cells(4.02, 0.02) % 4 != 0
    lockstep(oracle, cfg, make_scans(1.0, 720, 1.0, 3))


def test_step_parity_robot_near_border(oracle):
    """Map much smaller than the room: rays leave the grid, many endpoints are invalid."""
    cfg = GridMapSlamConfig(position=(-0.3, -0.5), width=1.0, height=1.0, resolution=0.02, n_particles=8)
    lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 4))


@COPY_MODES
def test_step_parity_rotated_rows_wrap_around(oracle, copy_flags):
    """Power-of-two grid (rows of a slot are stored rotated so that extents start on a DRAM page) with
    the robot a few cells from the left border: informed extents start at column 0, grow, and their
    physical image wraps around the row; particles are re-scattered so that slots change tenants."""
    cfg = GridMapSlamConfig(position=(-0.4, -2.56), width=5.12, height=5.12, resolution=0.02, n_particles=24)
    assert S.grid_cells(5.12, 0.02) == 256
    scans = make_scans(1.0, 360, 1.0, 6)
    rng = np.random.default_rng(5)

    def scatter(step, gpu, osl):
        if step in (2, 4):
            xyt = np.column_stack([rng.uniform(-0.3, 3.5, 24), rng.uniform(-2.0, 2.0, 24),
                                   rng.uniform(-np.pi, np.pi, 24)]).astype(np.float32)
            gpu.set_poses(xyt); osl.set_poses(xyt)

    errs = lockstep(oracle, cfg, scans, flags=copy_flags, pre_step=scatter)
    print(errs[-1])


@pytest.mark.parametrize("flags", [0, _lib.FLAG_FULL_GRID_COPY, _lib.FLAG_GENERIC_RAY_KERNEL | _lib.FLAG_UPDATE_ALL_PARTICLES],
                         ids=["extent-copy", "whole-slot-copy", "generic-strict"])
def test_windowed_slots_parity(oracle, flags):
    """1024 x 1024 logical grid, 256 x 256 slots (1/16 of the memory): each slot holds a torus window of
    the grid, enough for one particle's informed extent. Particles are re-scattered over the map, so
    slots change tenants whose extents lie in different parts of the grid (the windows alias) -- every
    read-out must still equal the oracle's full-grid result."""
    cfg = GridMapSlamConfig(position=(-10.24, -10.24), width=20.48, height=20.48, resolution=0.02, n_particles=20)
    assert S.grid_cells(20.48, 0.02) == 1024
    scans = make_scans(1.0, 360, 1.0, 6)      # 1 m range at 2 cm: extents of ~110 cells, growing
    rng = np.random.default_rng(9)

    def scatter(step, gpu, osl):
        if step == 0:     # while the maps are empty: a map that already spans two far-apart places cannot fit a window
            xyt = np.column_stack([rng.uniform(-9.0, 9.0, 20), rng.uniform(-9.0, 9.0, 20),
                                   rng.uniform(-np.pi, np.pi, 20)]).astype(np.float32)
            gpu.set_poses(xyt); osl.set_poses(xyt)

    errs = lockstep(oracle, cfg, scans, flags=flags, pre_step=scatter, slot_cells=256, particles=range(0, 20, 3))
    print(errs[-1])


def test_windowed_slots_memory_and_overflow():
    cfg = GridMapSlamConfig(position=(-10.24, -10.24), width=20.48, height=20.48, resolution=0.02, n_particles=8)
    (obs, odo), = make_scans(1.0, 360, 1.0, 1)
    with GridMapSlam(cfg, GpuPlacement(slot_cells=256)) as g:
        assert g.stats()["bytes_per_grid"] == 256 * 256 * 4          # not 1024 * 1024 * 4
        g.update(obs, odo)
        x0, y0, x1, y1 = g.map_extent()
        assert 0 < x1 - x0 <= 256 and 0 < y1 - y0 <= 256
        # send the particles far away: old extent + new extent no longer fit one 256 x 256 window
        g.set_poses(np.tile(np.array([[6.0, 6.0, 0.0]], np.float32), (8, 1)))
        with pytest.raises(_lib.SlamrsGpuError) as e:
            g.update(obs, odo)
        assert e.value.code == _lib.E_WINDOW and g.stats()["window_overflow"] > 0
    with pytest.raises(_lib.SlamrsGpuError):
        GridMapSlam(cfg, GpuPlacement(slot_cells=300))               # not a power of two
    # set_cells / cells round trip through a windowed slot, extent anywhere in the grid
    img = np.zeros((1024, 1024), np.uint32)
    img[700:900, 800:1000] = np.random.default_rng(1).integers(1, 5, (200, 200)).astype(np.uint32)
    with GridMapSlam(cfg, GpuPlacement(slot_cells=256)) as g:
        g.set_cells(2, img)
        assert np.array_equal(g.cells(2).reshape(1024, 1024), img)
        big = img.copy(); big[10, 10] = 1
        with pytest.raises(_lib.SlamrsGpuError) as e:
            g.set_cells(3, big)                                      # extent 990 x 890 does not fit
        assert e.value.code == _lib.E_WINDOW


def test_c5_shape(oracle):
    """configs[4] shape at reduced particle count: 720 beams at 0.5 degree spacing, 2048 x 2048 grid at
    5 cm (102.4 m), 6 m range, global-localisation-style uniform initial poses over the 20 m room."""
    cfg = GridMapSlamConfig(position=(-51.2, -51.2), width=102.4, height=102.4, resolution=0.05, n_particles=20)
    assert S.grid_cells(102.4, 0.05) == 2048
    scans = make_scans(10.0, 720, 6.0, 3)
    rng = np.random.default_rng(42)
    init = np.column_stack([rng.uniform(-9.0, 9.0, 20), rng.uniform(-9.0, 9.0, 20),
                            rng.uniform(-np.pi, np.pi, 20)]).astype(np.float32)

    def uniform_init(step, gpu, osl):
        if step == 0:
            gpu.set_poses(init); osl.set_poses(init)

    errs = lockstep(oracle, cfg, scans, particles=[0, 7, 19], pre_step=uniform_init)
    print(errs[-1])
    # the same with 512 x 512 slots: 1/16 of the memory per particle, which is what makes the full
    # configuration (32,768 particles per GPU) fit
    errs = lockstep(oracle, cfg, scans, particles=[0, 7, 19], pre_step=uniform_init, slot_cells=512)
    print(errs[-1])


@pytest.mark.parametrize("seed", range(32))
def test_step_parity_randomised_configurations(oracle, seed):
    """Random grid sizes (power-of-two and not, multiples of 8 and not), resolutions, beam counts,
    ranges, particle counts, kernel / copy modes and start poses (some next to a border, some
    re-scattered mid-run): every step in lockstep with the oracle. compute-sanitizer is not available
    on the GPU pool; out-of-bounds or stale-extent bugs show up here as a counter mismatch."""
    rng = np.random.default_rng(1000 + seed)
    cells = int(rng.choice([64, 96, 100, 128, 200, 256, 250, 512]))
    res = float(rng.choice([0.02, 0.04, 0.05]))
    width = cells * res
    n = int(rng.choice([1, 3, 16, 40]))
    beams = int(rng.choice([45, 180, 360, 720]))
    scene = float(rng.choice([1.0, 2.0, 5.0]))
    rng_m = float(rng.choice([0.5, 1.0, 3.0, 6.0]))
    flags = int(rng.choice([0, 0, _lib.FLAG_EAGER_COPY, _lib.FLAG_GENERIC_RAY_KERNEL, _lib.FLAG_FULL_GRID_COPY,
                            _lib.FLAG_UPDATE_ALL_PARTICLES, _lib.FLAG_EAGER_COPY | _lib.FLAG_UPDATE_ALL_PARTICLES,
                            _lib.FLAG_GENERIC_RAY_KERNEL | _lib.FLAG_UPDATE_ALL_PARTICLES]))
    off = rng.uniform(-0.45, 0.45, 2) * width          # where the robot starts inside the map
    cfg = GridMapSlamConfig(position=(float(-width / 2 + off[0]), float(-width / 2 + off[1])), width=width, height=width,
                            resolution=res, n_particles=n)
    assert S.grid_cells(width, res) in (cells, cells + 1)
    scans = make_scans(scene, beams, rng_m, 4)

    def scatter(step, gpu, osl):
        if step == 2 and n > 1:
            xyt = np.column_stack([rng.uniform(-0.4, 0.4, n) * width - off[0], rng.uniform(-0.4, 0.4, n) * width - off[1],
                                   rng.uniform(-np.pi, np.pi, n)]).astype(np.float32)
            gpu.set_poses(xyt); osl.set_poses(xyt)

    lockstep(oracle, cfg, scans, flags=flags, pre_step=scatter, particles=range(0, n, max(1, n // 6)))


def test_pose_outside_grid_emits_nothing(oracle):
    cfg = GridMapSlamConfig(position=(5.0, 5.0), width=1.0, height=1.0, resolution=0.02, n_particles=4)
    lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 2))


def test_empty_and_all_invalid_scans(oracle):
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=6)
    (obs, odo), = make_scans(1.0, 360, 1.0, 1)
    empty = Observation(0, angle=[], distance=[], valid=[])
    invalid = Observation(0, angle=obs.angle, distance=np.full_like(obs.distance, 1.0), valid=np.zeros(len(obs), bool))
    lockstep(oracle, cfg, [(empty, odo), (invalid, odo), (obs, odo), (empty, odo)])


def test_single_particle(oracle):
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=1)
    lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 3))


def test_many_particles_indices_bit_exact(oracle):
    """4096 particles on a small grid: stresses the scan / bisection / planner at scale."""
    cfg = GridMapSlamConfig(position=(-1.28, -1.28), width=2.56, height=2.56, resolution=0.04, n_particles=4096)
    errs = lockstep(oracle, cfg, make_scans(1.0, 360, 1.0, 5), particles=[0, 1, 777, 4095])
    print(errs[-1])


def test_destroy_create_cycle():
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=16)
    (obs, odo), = make_scans(1.0, 360, 1.0, 1)
    poses = []
    for _ in range(3):
        with GridMapSlam(cfg) as g:
            g.update(obs, odo)
            poses.append(g.poses().copy())
    assert np.array_equal(poses[0], poses[1]) and np.array_equal(poses[1], poses[2])


def test_invalid_arguments_are_errors_not_crashes():
    with pytest.raises(ValueError):
        GridMapSlam(GridMapSlamConfig(n_particles=0))
    with pytest.raises(_lib.SlamrsGpuError) as e:
        GridMapSlam(GridMapSlamConfig(position=(0, 0), width=4.0, height=2.0, resolution=0.02, n_particles=4))
    assert e.value.code == _lib.E_INVALID_ARG
    with GridMapSlam(GridMapSlamConfig(n_particles=4), GpuPlacement(rng_mode=_lib.RNG_CALLER)) as g:
        (obs, odo), = make_scans(1.0, 360, 1.0, 1)
        with pytest.raises(_lib.SlamrsGpuError):
            g.update(obs, odo)  # draws missing in CALLER mode


@pytest.mark.parametrize("copy_flags", [0, _lib.FLAG_EAGER_COPY], ids=["deferred", "eager"])
def test_resampling_conserves_grids_at_scale(copy_flags):
    """configs[1] shape (1,024 particles, 512^2 grid): size-independent properties of one step --
    every new particle's grid equals its source's post-update grid, indices are sorted,
    copies + distinct survivors == N, weights normalise to 1."""
    cfg = GridMapSlamConfig(position=(-12.8, -12.8), width=25.6, height=25.6, resolution=0.05, n_particles=1024)
    scans = make_scans(5.0, 360, 6.0, 3)
    with GridMapSlam(cfg, GpuPlacement(flags=copy_flags)) as g:
        for obs, odo in scans[:2]:
            g.update(obs, odo)
        probe = [0, 1, 2, 511, 1023]
        # grids before the third step, for a handful of sources we will look up afterwards
        g.update(*scans[2])
        idx = g.resample_indices().astype(np.int64)
        w, raw = g.weights()
        st = g.stats()
        assert np.all(np.diff(idx) >= 0)
        assert abs(w.sum() - 1.0) < 1e-12
        if copy_flags & _lib.FLAG_EAGER_COPY:
            assert st["grids_copied"] + st["distinct_sources"] == 1024
        else:   # a clone is copied only when it is about to be written
            assert st["grids_copied"] <= st["particles_integrated"] == st["distinct_sources"]
        assert st["distinct_sources"] == len(np.unique(idx))
        # duplicates of one source hold identical grids
        dup = np.nonzero(np.diff(idx) == 0)[0]
        for m in dup[:: max(1, len(dup) // 6)][:6]:
            assert np.array_equal(g.cells(int(m)), g.cells(int(m) + 1))
        assert st["counter_saturated"] == 0
