// Links the prebuilt libzkgpu.so (built by `python -m zkmips_b200.build`, nvcc -gencode arch=compute_100a,code=sm_100a).
// ZKGPU_LIB_DIR points at the directory holding it; the reference builds its own native code with `cc` in build.rs
// (crates/recursion/core/build.rs:193-196) -- here the CUDA build is kept outside cargo because it needs nvcc.
fn main() {
    let dir = std::env::var("ZKGPU_LIB_DIR").unwrap_or_else(|_| "../../zkmips_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=ZKGPU_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/zkgpu.h");
}
