// Links libminarrow_b200.so (built by `make -C minarrow_b200/csrc`: nvcc -gencode arch=compute_100a,code=sm_100a).
// Mirrors how the reference's own build.rs compiles and links its C helper (build.rs:28-34).
fn main() {
    let dir = std::env::var("MINARROW_B200_LIB_DIR")
        .expect("set MINARROW_B200_LIB_DIR to the directory that holds libminarrow_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=minarrow_b200");
    println!("cargo:rerun-if-env-changed=MINARROW_B200_LIB_DIR");
}
