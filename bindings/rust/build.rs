// Replaces the reference's build.rs:29-48,76 (cc-compile c_binder.cpp + link gomp/hdf5): nothing is compiled here, the
// crate links the prebuilt shared library. CLANN_B200_LIB_DIR = directory holding libclann_b200.so
// (clann_b200/lib in this repository after `python -m clann_b200._build`).
fn main() {
    let dir = std::env::var("CLANN_B200_LIB_DIR").expect("set CLANN_B200_LIB_DIR to the directory of libclann_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=clann_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CLANN_B200_LIB_DIR");
}
