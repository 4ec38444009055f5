// Links libresvg_b200.so (built by `make lib` at the repository root).
fn main() {
    let dir = std::env::var("RESVG_B200_LIB_DIR").unwrap_or_else(|_| "../../resvg_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=resvg_b200");
    println!("cargo:rerun-if-env-changed=RESVG_B200_LIB_DIR");
}
