"""The benchmark configurations of BASELINE.json, as data."""
from __future__ import annotations

from dataclasses import dataclass

from .simulator import Simulator, reference_scene
from .slam import GridMapSlamConfig


@dataclass(frozen=True)
class Workload:
    name: str
    n_particles: int        # per GPU for the weak-scaling bench (C4 = C3's shard on every GPU)
    n_beams: int
    grid: int               # cells per side
    resolution: float
    scene_scale: float
    scanner_range: float
    wheel_base: float
    speed_left: float = 0.08
    speed_right: float = 0.10
    update_period: float = 1.0
    slot_cells: int = 0          # windowed grid slots (GpuPlacement.slot_cells); 0 = whole-grid slots
    uniform_init: bool = False   # global-localisation-style start: poses uniform over the room

    @property
    def width(self) -> float:
        return self.grid * self.resolution

    def slam_config(self, n_particles=None) -> GridMapSlamConfig:
        w = self.width
        return GridMapSlamConfig(position=(-w / 2.0, -w / 2.0), width=w, height=w, resolution=self.resolution,
                                 n_particles=n_particles or self.n_particles)

    def simulator(self) -> Simulator:
        return Simulator(reference_scene(self.scene_scale), n_beams=self.n_beams, scanner_range=self.scanner_range,
                         wheel_base=self.wheel_base, update_period=self.update_period)


WORKLOADS = {
    # configs[0]: the shipped preset (config/grid_slam.yaml), 30 particles per BASELINE.json
    "c1": Workload("c1_30x360_200", 30, 360, 200, 0.02, 1.0, 1.0, 0.1),
    # configs[1]: 1,024 particles x 360 beams, 512^2 grid at 5 cm
    "c2": Workload("c2_1024x360_512", 1024, 360, 512, 0.05, 5.0, 6.0, 0.1),
    # configs[2]: 8,192 particles x 360 beams, 1024^2 grid (grid-copy-bound resampling)
    "c3": Workload("c3_8192x360_1024", 8192, 360, 1024, 0.05, 10.0, 6.0, 0.1),
    # configs[3] is c3's shard on each of 2/4/8 GPUs (65,536 particles at 8 GPUs)
    # configs[4]: 262,144 particles x 720 beams, 2048^2 grid on 8 GPUs = 32,768 per GPU. A whole 2048^2
    # grid per particle would need 512 GiB per GPU; 512 x 512 windowed slots (1 MiB) hold the informed
    # extent of a 6 m lidar with room to grow. Uniform initial poses over the 20 m room.
    "c5": Workload("c5_32768x720_2048", 32768, 720, 2048, 0.05, 10.0, 6.0, 0.1, slot_cells=512, uniform_init=True),
}


def uniform_poses(wl: Workload, first: int, count: int, seed: int = 0x5EED5A11):
    """Start poses for `uniform_init` workloads: x, y uniform over the scaled room, theta over [-pi, pi);
    a function of the GLOBAL particle index, so every sharding gives the same population."""
    import numpy as np
    half = 0.9 * wl.scene_scale          # the reference scene's outer rectangle is 2 m x 2 m, scaled
    out = np.empty((count, 3), np.float32)
    for i in range(count):
        rng = np.random.default_rng([seed, first + i])
        out[i] = (rng.uniform(-half, half), rng.uniform(-half, half), rng.uniform(-np.pi, np.pi))
    return out
