// Links the prebuilt C-ABI library. SLAMRS_GPU_LIB_DIR points at the directory that holds
// libslamrs_gpu.so (built by `python -m slamrs_b200.build`).
fn main() {
    let dir = std::env::var("SLAMRS_GPU_LIB_DIR").unwrap_or_else(|_| "../../slamrs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=slamrs_gpu");
    println!("cargo:rerun-if-env-changed=SLAMRS_GPU_LIB_DIR");
}
