// Link against libpcs.so.  PCS_LIB_DIR = the directory `python -m plonky2_demo_b200.build` left the library in
// (plonky2_demo_b200/ of this repository); the same directory must be on LD_LIBRARY_PATH at run time.
fn main() {
    if let Ok(dir) = std::env::var("PCS_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=pcs");
    println!("cargo:rerun-if-env-changed=PCS_LIB_DIR");
}
