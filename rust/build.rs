// Links libbjj_cuda.so (built by babyjubjub-rs_b200/build.py with nvcc for sm_100a).
fn main() {
    let dir = std::env::var("BJJ_CUDA_LIB_DIR")
        .expect("set BJJ_CUDA_LIB_DIR to the directory that holds libbjj_cuda.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bjj_cuda");
    println!("cargo:rerun-if-env-changed=BJJ_CUDA_LIB_DIR");
}
