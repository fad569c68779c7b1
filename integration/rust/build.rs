// Links libneurokmer.so (built by `python -m neurokmer_b200.build`).  NOT compiled in this repository's image.
fn main() {
    let dir = std::env::var("NEUROKMER_LIB_DIR").expect("set NEUROKMER_LIB_DIR to the directory holding libneurokmer.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=neurokmer");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
