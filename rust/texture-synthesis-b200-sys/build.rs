// Links the prebuilt CUDA library (python texture-synthesis_b200/build.py -> libtsb200.so, nvcc sm_100a).
fn main() {
    let dir = std::env::var("TSB200_LIB_DIR").expect("set TSB200_LIB_DIR to the directory holding libtsb200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tsb200");
    println!("cargo:rerun-if-env-changed=TSB200_LIB_DIR");
}
