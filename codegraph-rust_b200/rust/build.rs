// Links the prebuilt C-ABI library.  CGVEC_B200_LIB_DIR points at the directory holding libcgvec_b200.so.
fn main() {
    let dir = std::env::var("CGVEC_B200_LIB_DIR").unwrap_or_else(|_| "../".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cgvec_b200");
    println!("cargo:rerun-if-env-changed=CGVEC_B200_LIB_DIR");
}
