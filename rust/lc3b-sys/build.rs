// Links the in-tree shared library built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("LC3B_LIB_DIR").unwrap_or_else(|_| "../../lc3_codec_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=lc3b");
    println!("cargo:rerun-if-env-changed=LC3B_LIB_DIR");
}
