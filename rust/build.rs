// Builds (or finds) libstreamz_b200.so and links it.  Replaces the reference's vestigial build.rs
// (streamz-rs/build.rs:1-66 only writes an unused train_files.rs).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("streamz_b200").join("csrc");
    let lib_dir = root.join("streamz_b200").join("lib");
    if env::var_os("STREAMZ_B200_PREBUILT").is_none() {
        // nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see streamz_b200/csrc/Makefile)
        let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make not found");
        assert!(status.success(), "building libstreamz_b200.so failed");
    }
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=streamz_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    for f in ["capi.cu", "frontend.cu", "mlp.cu", "comm.cu", "formats.cu", "gemm_tc.cuh", "fft_math.cuh", "tables.hpp"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("streamz_b200.h").display());
}
