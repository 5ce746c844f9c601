// build.rs -- the cargo build script of the north star: compiles the CUDA library of this repository with nvcc
// for sm_100a and links the crate against it.
//
//   KZG_B200_CSRC   directory holding the CUDA sources (default: ../kzg_rust_b200/csrc next to this crate)
//   KZG_B200_LIB_DIR  skip the build and link a prebuilt libkzg_b200.so from this directory instead
//   NVCC            compiler driver (default: nvcc)
//
// The translation units are the ones kzg_rust_b200/build.py compiles (keep the two lists in step).
use std::env;
use std::path::PathBuf;
use std::process::Command;

const SOURCES: [&str; 6] = ["kzg_b200.cu", "msm.cu", "g1ops.cu", "frops.cu", "host_pairing.cpp", "host_sha256.cpp"];
const NVCC_FLAGS: [&str; 9] = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "--expt-relaxed-constexpr",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-shared",
];

fn main() {
    println!("cargo:rerun-if-env-changed=KZG_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=KZG_B200_CSRC");
    println!("cargo:rerun-if-env-changed=NVCC");
    if let Ok(dir) = env::var("KZG_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=kzg_b200");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
    let csrc = env::var("KZG_B200_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("..").join("kzg_rust_b200").join("csrc"));
    let out_dir = PathBuf::from(env::var("OUT_DIR").expect("OUT_DIR"));
    let lib = out_dir.join("libkzg_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let mut cmd = Command::new(&nvcc);
    cmd.args(NVCC_FLAGS).arg("-o").arg(&lib);
    for src in SOURCES {
        cmd.arg(csrc.join(src));
    }
    let status = cmd
        .status()
        .unwrap_or_else(|e| panic!("could not run {}: {} (there is no CPU fallback: the crate needs the CUDA library)", nvcc, e));
    assert!(status.success(), "nvcc failed on {}", csrc.display());
    println!("cargo:rustc-link-search=native={}", out_dir.display());
    println!("cargo:rustc-link-lib=dylib=kzg_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("..").join("include").join("kzg_b200.h").display());
}
