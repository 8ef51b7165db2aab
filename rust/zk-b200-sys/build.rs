// SOURCE ONLY (see Cargo.toml).  Links libzk_b200.so from $ZK_B200_LIB_DIR (default: ../../zk_b200).
fn main() {
    let dir = std::env::var("ZK_B200_LIB_DIR").unwrap_or_else(|_| "../../zk_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zk_b200");
    println!("cargo:rerun-if-env-changed=ZK_B200_LIB_DIR");
}
