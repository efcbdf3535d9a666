"""The Rust drop-in (rust/drop-in/slam.rs + rust/slam-gpu-sys) cannot be compiled here (no Rust toolchain in
the image), so it is checked mechanically: the shim keeps the exact public surface the unchanged node uses
(slamrs/slam/src/grid/slam.rs:18-25, 28, 46, 77, 83, 90; node.rs:12-15, 53-57), returns the reference's own
`super::map::GridData<Probability>`, checks every status code, and every FFI item it touches is declared in
the -sys crate with the field / parameter lists of include/slamrs_gpu.h."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = open(os.path.join(ROOT, "rust", "drop-in", "slam.rs")).read()
SYS = open(os.path.join(ROOT, "rust", "slam-gpu-sys", "src", "lib.rs")).read()
HDR = open(os.path.join(ROOT, "include", "slamrs_gpu.h")).read()


def _norm(s):
    return re.sub(r"\s+", " ", s).strip()


def test_shim_keeps_the_reference_signatures():
    code = _norm(SHIM)
    for sig in ("pub fn new(config: &GridMapSlamConfig) -> Self",                 # slam.rs:28
                "pub fn update(&mut self, z: &Observation, u: Odometry)",         # slam.rs:46
                "pub fn estimated_pose(&self) -> Pose",                           # slam.rs:77
                "pub fn estimated_likelihood(&self) -> GridData<Probability>",    # slam.rs:83
                "pub fn map_position(&self) -> Vector2<f32>"):                    # slam.rs:90
        assert sig in code, sig
    # slam.rs:18-25: the serde shape of the YAML block (n_particles stays private)
    m = re.search(r"#\[derive\(Deserialize, Clone\)\] pub struct GridMapSlamConfig \{(.*?)\}", code)
    assert m and [f.strip() for f in m.group(1).split(",") if f.strip()] == [
        "pub position: Vector2<f32>", "pub width: f32", "pub height: f32", "pub resolution: f32", "n_particles: usize"]
    # the published map is the reference's own grid type (node.rs:12-15, 68-72; visualize.rs:248-252 reads Cell.column / .row)
    assert "use super::map::GridData;" in SHIM and "struct GridData" not in SHIM and "GridData::from_vec(" in SHIM
    patch = open(os.path.join(ROOT, "rust", "drop-in", "map_from_vec.patch")).read()
    assert "pub(crate) fn from_vec(size: Vector2<usize>, data: Vec<T>) -> Self" in patch


def test_shim_checks_every_status_code():
    calls = re.findall(r"(let (\w+) = )?unsafe \{\s*sys::(slamrs_gpu_\w+)\(", SHIM) + \
            re.findall(r"(let (\w+) = )unsafe \{\s*\n\s*sys::(slamrs_gpu_\w+)\(", SHIM)
    seen = {name for _, _, name in calls}
    assert {"slamrs_gpu_grid_cells", "slamrs_gpu_create", "slamrs_gpu_update", "slamrs_gpu_pose",
            "slamrs_gpu_map_probability", "slamrs_gpu_destroy"} <= seen | {"slamrs_gpu_destroy"}
    for assign, var, name in calls:
        if name in ("slamrs_gpu_destroy", "slamrs_gpu_last_error"):
            continue
        assert var, f"{name}: status code dropped"
        assert re.search(rf"{var} [!=]= sys::SLAMRS_OK", SHIM), f"{name}: {var} never compared with SLAMRS_OK"


def _c_fields(struct):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), HDR, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for d in body.split(";"):
        if d.strip():   # `float pos_x, pos_y` declares two fields
            out += [re.sub(r"\[.*\]", "", part.strip().split()[-1]) for part in d.split(",")]
    return out


def _rust_fields(struct, text=SYS):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, text, flags=re.S).group(1)
    return re.findall(r"pub (\w+):", body)


def test_sys_crate_mirrors_the_header():
    assert _rust_fields("slamrs_gpu_config") == _c_fields("slamrs_gpu_config")
    assert _rust_fields("slamrs_gpu_stats") == _c_fields("slamrs_gpu_stats")
    # every FFI function the shim calls is declared in the -sys crate with the header's parameter count
    for name in set(re.findall(r"sys::(slamrs_gpu_\w+)\(", SHIM)):
        rust = re.search(r"pub fn %s\((.*?)\)" % name, SYS, flags=re.S)
        c = re.search(r"\b%s\((.*?)\);" % name, HDR, flags=re.S)
        assert rust and c, name
        n_rust = len([a for a in rust.group(1).split(",") if a.strip()])
        n_c = len([a for a in re.sub(r"/\*.*?\*/", "", c.group(1), flags=re.S).split(",") if a.strip() and a.strip() != "void"])
        assert n_rust == n_c, (name, n_rust, n_c)
    # the config literal in the shim names every field exactly once
    lit = re.search(r"sys::slamrs_gpu_config \{(.*?)\n        \};", SHIM, flags=re.S).group(1)
    assert re.findall(r"^\s{12}(\w+):", lit, flags=re.M) == _c_fields("slamrs_gpu_config")
