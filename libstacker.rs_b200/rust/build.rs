// Builds the CUDA static library with nvcc for sm_100a and links it (plus cudart) into the crate.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../csrc");
    let include = manifest.join("../../include");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let obj = out.join("stacker_cuda.o");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"])
        .arg("-I").arg(&include)
        .arg("-c").arg(csrc.join("stacker_cuda.cu"))
        .arg("-o").arg(&obj)
        .status()
        .expect("nvcc not found: set NVCC or put the CUDA 12.9+ toolkit on PATH");
    assert!(status.success(), "nvcc failed");
    let lib = out.join("libstacker_cuda.a");
    assert!(Command::new("ar").arg("rcs").arg(&lib).arg(&obj).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=stacker_cuda");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for f in ["stacker_cuda.cu", "common.cuh", "prep.cuh", "ecc_iter.cuh", "ecc_iter_v2.cuh", "warp_acc.cuh", "tenengrad.cuh",
              "resize_area.cuh", "peer_reduce.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", include.join("stacker_cuda.h").display());
}
