// UNCOMPILED (no Rust toolchain in the build image).  Link against the in-tree libtfs_cuda.so.
fn main() {
    let dir = std::env::var("TFS_LIB_DIR").unwrap_or_else(|_| "../../terminal_fluid_sim_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tfs_cuda");
    println!("cargo:rerun-if-env-changed=TFS_LIB_DIR");
}
