// TSS_LIB_DIR = directory holding libtss.so (timberborn_support_solver_b200/ in this repository)
fn main() {
    let dir = std::env::var("TSS_LIB_DIR").expect("set TSS_LIB_DIR to the directory that holds libtss.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=tss");
    println!("cargo:rerun-if-env-changed=TSS_LIB_DIR");
}
