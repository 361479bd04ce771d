// Link against librbrt_gpu.so; RBRT_GPU_LIB_DIR points at the directory that holds it (rbrt_b200/).
fn main() {
    if let Ok(dir) = std::env::var("RBRT_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=rbrt_gpu");
    println!("cargo:rerun-if-env-changed=RBRT_GPU_LIB_DIR");
}
