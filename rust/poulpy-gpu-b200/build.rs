// Links libpoulpy_b200.so (built by poulpy_b200/csrc/build.sh); POULPY_B200_LIB_DIR points at the directory that holds it.
fn main() {
    let dir = std::env::var("POULPY_B200_LIB_DIR").unwrap_or_else(|_| "../../poulpy_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=poulpy_b200");
    println!("cargo:rerun-if-env-changed=POULPY_B200_LIB_DIR");
}
