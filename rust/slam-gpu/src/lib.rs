//! Drop-in `GridMapSlam` backed by libslamrs_gpu.so. Mirrors slamrs/slam/src/grid/slam.rs:13-97.
use std::ffi::CStr;
use std::ptr;

use common::math::Probability;
use common::robot::{Observation, Odometry, Pose};
use nalgebra::Vector2;
use serde::Deserialize;
use slam_gpu_sys as sys;

/// Same fields, same serde shape as the reference's `GridMapSlamConfig` (slam.rs:18-25).
#[derive(Deserialize, Clone)]
pub struct GridMapSlamConfig {
    pub position: Vector2<f32>,
    pub width: f32,
    pub height: f32,
    pub resolution: f32,
    n_particles: usize,
}

/// `GridData<Probability>` as the visualizer consumes it (map.rs:181-264): `size` + flat data.
pub struct GridData<T> {
    size: Vector2<usize>,
    data: Vec<T>,
}
impl<T> GridData<T> {
    pub fn size(&self) -> Vector2<usize> { self.size }
    pub fn iter_cells(&self) -> impl Iterator<Item = ((usize, usize), &T)> {
        let sy = self.size.y;
        self.data.iter().enumerate().map(move |(i, v)| ((i / sy, i % sy), v)) // (row, column)
    }
}

pub struct GridMapSlam {
    h: *mut sys::slamrs_gpu_handle,
    position: Vector2<f32>,
    grid: Vector2<usize>,
    scratch: (Vec<f32>, Vec<f32>, Vec<u8>),
}
// driven from one thread at a time (the GUI thread, baseui/src/app.rs:138-140)
unsafe impl Send for GridMapSlam {}

fn last_error(h: *const sys::slamrs_gpu_handle) -> String {
    unsafe { CStr::from_ptr(sys::slamrs_gpu_last_error(h)).to_string_lossy().into_owned() }
}

impl GridMapSlam {
    pub fn new(config: &GridMapSlamConfig) -> Self {
        assert!(config.n_particles > 0, "Must have at least one particle"); // particle.rs:16
        let (mut gw, mut gh) = (0u32, 0u32);
        unsafe {
            sys::slamrs_gpu_grid_cells(config.width, config.resolution, &mut gw);
            sys::slamrs_gpu_grid_cells(config.height, config.resolution, &mut gh);
        }
        let cfg = sys::slamrs_gpu_config {
            struct_size: std::mem::size_of::<sys::slamrs_gpu_config>() as u32,
            abi_version: sys::SLAMRS_GPU_ABI_VERSION,
            pos_x: config.position.x,
            pos_y: config.position.y,
            resolution: config.resolution,
            grid_w: gw,
            grid_h: gh,
            n_particles: config.n_particles as u64,
            seed: std::env::var("SLAMRS_SEED").ok().and_then(|s| s.parse().ok()).unwrap_or(0x5EED5A11),
            rng_mode: sys::SLAMRS_RNG_SHARED_STREAM,
            device: std::env::var("SLAMRS_GPU_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(-1),
            rank: 0,
            world_size: 1,
            spare_slots: 0,
            flags: 0,
            slot_cells: std::env::var("SLAMRS_SLOT_CELLS").ok().and_then(|s| s.parse().ok()).unwrap_or(0),
            reserved0: 0,
            nccl_id: [0; sys::SLAMRS_NCCL_ID_BYTES],
        };
        let mut h = ptr::null_mut();
        let rc = unsafe { sys::slamrs_gpu_create(&cfg, &mut h) };
        // the reference constructor is infallible; a missing GPU is a configuration error
        assert!(rc == sys::SLAMRS_OK, "slamrs_gpu_create failed ({rc}): {}", last_error(ptr::null()));
        GridMapSlam { h, position: config.position, grid: Vector2::new(gw as usize, gh as usize), scratch: Default::default() }
    }

    #[tracing::instrument(skip_all)]
    pub fn update(&mut self, z: &Observation, u: Odometry) {
        let (a, d, v) = &mut self.scratch;
        a.clear(); d.clear(); v.clear();
        for m in &z.measurements {
            a.push(m.angle as f32);     // the casts of map.rs:76-77, 121-122
            d.push(m.distance as f32);
            v.push(m.valid as u8);
        }
        let rc = unsafe {
            sys::slamrs_gpu_update(self.h, a.as_ptr(), d.as_ptr(), v.as_ptr(), a.len() as u32,
                u.distance_left, u.distance_right, u.wheel_distance, ptr::null(), ptr::null())
        };
        if rc != sys::SLAMRS_OK {
            // update() cannot return an error in the reference API: log and keep the last state
            tracing::error!("slamrs_gpu_update failed ({rc}): {}", last_error(self.h));
        }
    }

    pub fn estimated_pose(&self) -> Pose {
        let mut o = [0f32; 3];
        unsafe { sys::slamrs_gpu_pose(self.h, o.as_mut_ptr()) };
        Pose { x: o[0], y: o[1], theta: o[2] }
    }

    pub fn estimated_likelihood(&self) -> GridData<Probability> {
        let n = self.grid.x * self.grid.y;
        let mut raw = vec![0f64; n];
        unsafe { sys::slamrs_gpu_map_probability(self.h, raw.as_mut_ptr()) };
        GridData { size: self.grid, data: raw.into_iter().map(Probability::new_unchecked).collect() }
    }

    pub fn map_position(&self) -> Vector2<f32> { self.position }

    /// ParticleFilter::number_of_effective_particles (particle.rs:59-65) of the last update's
    /// normalised weights, before resampling reset them to 1/N.
    pub fn number_of_effective_particles(&self) -> f64 {
        let mut v = 0f64;
        unsafe { sys::slamrs_gpu_effective_particles(self.h, &mut v) };
        v
    }

    /// The informed window of the published map as f32 grey levels: `(x0, y0, x1, y1)` in cells and
    /// `(y1 - y0) * (x1 - x0)` values, row-major. Every cell outside the window is exactly 0.5.
    /// What the visualizer needs (baseui/src/node/visualize.rs:245-256) at 1/100 of the bytes of
    /// `estimated_likelihood()`.
    pub fn estimated_likelihood_window(&self) -> ((usize, usize, usize, usize), Vec<f32>) {
        let mut e = [0i32; 4];
        unsafe { sys::slamrs_gpu_map_extent(self.h, e.as_mut_ptr()) };
        let (x0, y0, x1, y1) = (e[0].max(0) as usize, e[1].max(0) as usize, e[2].max(0) as usize, e[3].max(0) as usize);
        let mut data = vec![0f32; (x1 - x0) * (y1 - y0)];
        if !data.is_empty() {
            unsafe {
                sys::slamrs_gpu_map_window(self.h, sys::SLAMRS_MAP_F32, e[0], e[1], e[2], e[3],
                    data.as_mut_ptr() as *mut core::ffi::c_void)
            };
        }
        ((x0, y0, x1, y1), data)
    }
}

impl Drop for GridMapSlam {
    fn drop(&mut self) { unsafe { sys::slamrs_gpu_destroy(self.h) } }
}
