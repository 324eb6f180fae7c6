// Locates (or, with the `build-lib` feature, builds) libaccmsm.so and tells cargo to link it.
//   ACCMSM_LIB_DIR   directory holding libaccmsm.so (default: ../accumulation_b200 relative to this crate)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.parent().unwrap().to_path_buf();
    let lib_dir = env::var("ACCMSM_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| repo.join("accumulation_b200"));
    if env::var("CARGO_FEATURE_BUILD_LIB").is_ok() {
        // the repository's own recipe: nvcc -gencode arch=compute_100a,code=sm_100a ... -o accumulation_b200/libaccmsm.so
        let status = Command::new("make").current_dir(&repo).status().expect("failed to run make");
        assert!(status.success(), "make (libaccmsm.so) failed");
    }
    assert!(lib_dir.join("libaccmsm.so").exists(), "libaccmsm.so not found in {} (set ACCMSM_LIB_DIR or enable the build-lib feature)", lib_dir.display());
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=accmsm");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-env-changed=ACCMSM_LIB_DIR");
    println!("cargo:rerun-if-changed={}", repo.join("include/accmsm.h").display());
}
