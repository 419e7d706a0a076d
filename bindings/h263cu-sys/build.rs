// Links libh263cu.so.  H263CU_LIB_DIR names the directory that holds it (the in-tree build puts it in
// <repo>/h263_rs_b200/); without the variable the system search path is used.  The library links the CUDA runtime
// statically, so nothing else is needed at link time.
fn main() {
    println!("cargo:rerun-if-env-changed=H263CU_LIB_DIR");
    if let Ok(dir) = std::env::var("H263CU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=h263cu");
}
