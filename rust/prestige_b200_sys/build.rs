// Link against the prebuilt libprestige_b200.so (make -C prestige_b200/csrc).  PST_B200_LIB_DIR overrides the search dir.
fn main() {
    let dir = std::env::var("PST_B200_LIB_DIR").unwrap_or_else(|_| "../../prestige_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=prestige_b200");
    println!("cargo:rerun-if-env-changed=PST_B200_LIB_DIR");
}
