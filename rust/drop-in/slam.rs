//! REPLACEMENT for `slamrs/slam/src/grid/slam.rs` (the whole file): the same `GridMapSlam` /
//! `GridMapSlamConfig` the node uses (slam.rs:13-97), backed by libslamrs_gpu.so through the
//! `slam-gpu-sys` crate. It lives INSIDE the `slam` crate, next to the unchanged `node.rs`, so every
//! type the node names stays what it was:
//!
//!   * `node.rs:12-15` imports `super::map::GridData` and `super::slam::{GridMapSlam, GridMapSlamConfig}`;
//!   * `node.rs:53-57` stores `self.slam.estimated_likelihood()` in `GridMapMessage.data`, declared as
//!     `super::map::GridData<Probability>` (node.rs:68-72) -- so this file returns exactly that type;
//!   * `baseui/src/node/visualize.rs:248-252` iterates `data.iter_cells()` and reads `c.column` / `c.row` of
//!     `super::map::Cell` -- untouched, because the `GridData` is the reference's own.
//!
//! The one change outside this file: `GridData`'s fields are private to `map.rs`, so `map.rs` gets the
//! three-line constructor in `map_from_vec.patch` (`pub(crate) fn from_vec`). `Cargo.toml` gains the
//! `slam-gpu-sys` dependency (`Cargo.toml.patch`). `node.rs`, `GridMapSlamNodeConfig`, the YAML and the
//! pubsub topics are byte-for-byte the reference's.
//!
//! Not compiled in this repository (no Rust toolchain in the build image): checked by eye against the
//! reference sources named above and against include/slamrs_gpu.h; tests/test_rust_dropin.py checks the
//! signatures mechanically.
use std::ffi::CStr;
use std::ptr;

use common::math::Probability;
use common::robot::{Observation, Odometry, Pose};
use nalgebra::Vector2;
use serde::Deserialize;
use slam_gpu_sys as sys;

use super::map::GridData;

/// Same fields, same serde shape as the reference's `GridMapSlamConfig` (slam.rs:18-25).
#[derive(Deserialize, Clone)]
pub struct GridMapSlamConfig {
    pub position: Vector2<f32>,
    pub width: f32,
    pub height: f32,
    pub resolution: f32,
    n_particles: usize,
}

pub struct GridMapSlam {
    h: *mut sys::slamrs_gpu_handle,
    position: Vector2<f32>,
    grid: Vector2<usize>,
    scratch: (Vec<f32>, Vec<f32>, Vec<u8>),
    /// what the last successful read-outs returned: the reference API is infallible, so a failing call
    /// logs and repeats the last good value instead of inventing zeros
    last_pose: std::cell::Cell<Pose>,
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
        let rc_w = unsafe { sys::slamrs_gpu_grid_cells(config.width, config.resolution, &mut gw) };
        let rc_h = unsafe { sys::slamrs_gpu_grid_cells(config.height, config.resolution, &mut gh) };
        assert!(rc_w == sys::SLAMRS_OK && rc_h == sys::SLAMRS_OK, "grid size out of range");
        let env = |k: &str| std::env::var(k).ok();
        let cfg = sys::slamrs_gpu_config {
            struct_size: std::mem::size_of::<sys::slamrs_gpu_config>() as u32,
            abi_version: sys::SLAMRS_GPU_ABI_VERSION,
            pos_x: config.position.x,
            pos_y: config.position.y,
            resolution: config.resolution,
            grid_w: gw,
            grid_h: gh,
            n_particles: config.n_particles as u64,
            seed: env("SLAMRS_SEED").and_then(|s| s.parse().ok()).unwrap_or(0x5EED5A11),
            rng_mode: sys::SLAMRS_RNG_SHARED_STREAM,
            device: env("SLAMRS_GPU_DEVICE").and_then(|s| s.parse().ok()).unwrap_or(-1),
            rank: 0,
            world_size: 1,
            spare_slots: 0,
            flags: 0,
            slot_cells: env("SLAMRS_SLOT_CELLS").and_then(|s| s.parse().ok()).unwrap_or(0),
            resample_threshold: 0.0, // resample after every update, as slam.rs:74 does
            nccl_id: [0; sys::SLAMRS_NCCL_ID_BYTES],
        };
        let mut h = ptr::null_mut();
        let rc = unsafe { sys::slamrs_gpu_create(&cfg, &mut h) };
        // the reference constructor is infallible; a missing GPU is a configuration error
        assert!(rc == sys::SLAMRS_OK, "slamrs_gpu_create failed ({rc}): {}", last_error(ptr::null()));
        GridMapSlam {
            h,
            position: config.position,
            grid: Vector2::new(gw as usize, gh as usize),
            scratch: Default::default(),
            last_pose: std::cell::Cell::new(Pose::default()),
        }
    }

    #[tracing::instrument(skip_all)]
    pub fn update(&mut self, z: &Observation, u: Odometry) {
        let (a, d, v) = &mut self.scratch;
        a.clear();
        d.clear();
        v.clear();
        for m in &z.measurements {
            a.push(m.angle as f32); // the casts of map.rs:76-77, 121-122
            d.push(m.distance as f32);
            v.push(m.valid as u8);
        }
        let rc = unsafe {
            sys::slamrs_gpu_update(
                self.h, a.as_ptr(), d.as_ptr(), v.as_ptr(), a.len() as u32,
                u.distance_left, u.distance_right, u.wheel_distance, ptr::null(), ptr::null(),
            )
        };
        if rc != sys::SLAMRS_OK {
            // update() cannot return an error in the reference API: log and keep the last state
            tracing::error!("slamrs_gpu_update failed ({rc}): {}", last_error(self.h));
        }
    }

    pub fn estimated_pose(&self) -> Pose {
        let mut o = [0f32; 3];
        let rc = unsafe { sys::slamrs_gpu_pose(self.h, o.as_mut_ptr()) };
        if rc != sys::SLAMRS_OK {
            tracing::error!("slamrs_gpu_pose failed ({rc}): {}", last_error(self.h));
            return self.last_pose.get();
        }
        let p = Pose { x: o[0], y: o[1], theta: o[2] };
        self.last_pose.set(p);
        p
    }

    /// slam.rs:83-88. Returns the reference's own `GridData<Probability>` (map.rs:181-264).
    pub fn estimated_likelihood(&self) -> GridData<Probability> {
        let n = self.grid.x * self.grid.y;
        let mut raw = vec![0.5f64; n]; // the prior: what an empty map reads as
        let rc = unsafe { sys::slamrs_gpu_map_probability(self.h, raw.as_mut_ptr()) };
        if rc != sys::SLAMRS_OK {
            tracing::error!("slamrs_gpu_map_probability failed ({rc}): {}", last_error(self.h));
            raw.iter_mut().for_each(|p| *p = 0.5);
        }
        GridData::from_vec(self.grid, raw.into_iter().map(Probability::new_unchecked).collect())
    }

    /// slam.rs:90-96
    pub fn map_position(&self) -> Vector2<f32> {
        self.position
    }
}

impl Drop for GridMapSlam {
    fn drop(&mut self) {
        unsafe { sys::slamrs_gpu_destroy(self.h) }
    }
}
