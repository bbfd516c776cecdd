// Links libspfresh_b200.so.  SPFRESH_B200_LIB_DIR points at the directory that holds it (the
// `spfresh_b200/` directory of the B200 repo after `python -m spfresh_b200.build`).
fn main() {
    println!("cargo:rerun-if-env-changed=SPFRESH_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("SPFRESH_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=spfresh_b200");
}
