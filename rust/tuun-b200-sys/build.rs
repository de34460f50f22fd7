fn main() {
    // point TUUN_B200_LIB_DIR at the directory holding libtuun_b200.so
    if let Ok(dir) = std::env::var("TUUN_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=tuun_b200");
}
