// build.rs -- compiles the .cu library with nvcc for sm_100a and links it.
// NOTE: written but NOT compiled in the build image (no cargo/rustc there); see INTEGRATION.md.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let src = root.join("image_webp_b200/csrc/zw_capi.cu");
    let lib = out.join("libzenwebp_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("nvcc not found: there is no CPU fallback for this crate");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=zenwebp_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", root.join("image_webp_b200/csrc").display());
}
