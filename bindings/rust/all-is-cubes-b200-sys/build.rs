//! Links `libaicb200.so` (built in-tree by `__graft_entry__.build()`; static cudart, no other dependency than the
//! CUDA driver).  `AICB200_LIB_DIR` names the directory that holds it.
fn main() {
    let dir = std::env::var("AICB200_LIB_DIR").unwrap_or_else(|_| "../../../all-is-cubes_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=aicb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=AICB200_LIB_DIR");
}
