// Builds libsawfish_cuda.so with nvcc for sm_100a and links it.  No other architecture, no CPU fallback:
// if nvcc is missing the build fails.
// SAWFISH_CUDA_CSRC points at the kernel sources (this repository's sawfish_b200/csrc, which includes
// ../../include/sawfish_cuda.h); by default they are expected next to the crate as in this repository.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("SAWFISH_CUDA_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../sawfish_b200/csrc"));
    let lib = out.join("libsawfish_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC,-O3", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("sfc_lib.cu"))
        .status()
        .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e}"));
    assert!(status.success(), "nvcc failed (sm_100a build is mandatory; there is no fallback)");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=sawfish_cuda");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=NVCC");
    println!("cargo:rerun-if-env-changed=SAWFISH_CUDA_CSRC");
}
