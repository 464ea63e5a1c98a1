// UNVERIFIED (never compiled: no rustc in the build image).
// Builds libndstats_b200.so with the in-tree Makefile (nvcc, sm_100a) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make");
    assert!(status.success(), "building libndstats_b200.so failed (needs nvcc with sm_100a support)");
    println!("cargo:rustc-link-search=native={}", manifest.join("..").display());
    println!("cargo:rustc-link-lib=dylib=ndstats_b200");
    println!("cargo:rerun-if-changed=../csrc");
    println!("cargo:rerun-if-changed=../../include/ndstats_b200.h");
}
