// Links against the in-tree build of the CUDA library:
//   python -m lcpc_proof_of_storage_b200.build   ->  lcpc_proof_of_storage_b200/_lib/liblcpc_b200.so
fn main() {
    let dir = std::env::var("LCPC_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../../lcpc_proof_of_storage_b200/_lib", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=lcpc_b200");
    println!("cargo:rerun-if-env-changed=LCPC_B200_LIB_DIR");
}
