// Builds and links libminarrow_b200.so — the hand-written sm_100a CUDA kernels behind the C ABI (include/minarrow_b200.h).
// Mirrors how the reference's own build.rs compiles and links its C helper with `cc` (build.rs:28-34,113-154), except that
// the compiler is nvcc: `make -C minarrow_b200/csrc` runs
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true ... -shared
// (one translation unit per element type, built in parallel).  There is no other backend and no CPU fallback.
//
//   MINARROW_B200_LIB_DIR   use a prebuilt library in this directory instead of building
//   NVCC                    nvcc to use (default /usr/local/cuda/bin/nvcc, see the Makefile)
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=MINARROW_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");
    let dir = match std::env::var("MINARROW_B200_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            // rust/minarrow-b200/ -> repo root
            let root = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
            let csrc = root.join("minarrow_b200/csrc");
            for f in ["api.cu", "elementwise.cu", "reduce.cu", "bits.cu", "compare.cu", "concat.cu", "arrow_ffi.cu",
                      "ew_kernels.cuh", "reduce_kernels.cuh", "common.cuh", "shift_load.cuh", "divmagic.h", "fastmod.h",
                      "internal.h", "Makefile"] {
                println!("cargo:rerun-if-changed={}", csrc.join(f).display());
            }
            println!("cargo:rerun-if-changed={}", root.join("include/minarrow_b200.h").display());
            let jobs = std::thread::available_parallelism().map(|n| n.get()).unwrap_or(4).to_string();
            let status = Command::new("make").arg("-C").arg(&csrc).arg("-j").arg(&jobs).arg("-s")
                .status().expect("minarrow-b200: could not run `make` (needs make + nvcc 12.8 or newer for sm_100a)");
            assert!(status.success(), "minarrow-b200: building the CUDA kernels failed");
            root.join("minarrow_b200")
        }
    };
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=minarrow_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
