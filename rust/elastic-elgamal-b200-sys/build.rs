// Links against the in-tree CUDA library built by `python -m elastic_elgamal_b200.build`.
fn main() {
    let dir = std::env::var("EG_B200_LIB_DIR").unwrap_or_else(|_| "../../elastic_elgamal_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=eg_b200");
    println!("cargo:rerun-if-env-changed=EG_B200_LIB_DIR");
}
