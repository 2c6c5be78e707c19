// Links the prebuilt CUDA library (make -> rna_algos_b200/librna_algos_b200.so).  No bindgen: src/ffi.rs is a
// hand-written image of include/rna_algos_b200.h, checked at run time against rna_sizeof_*_tables().
fn main() {
    let dir = std::env::var("RNA_ALGOS_B200_LIB_DIR").unwrap_or_else(|_| "../../rna_algos_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rna_algos_b200");
    println!("cargo:rerun-if-env-changed=RNA_ALGOS_B200_LIB_DIR");
}
