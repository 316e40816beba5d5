// build.rs — compiles the CUDA sources of liboxg for sm_100a and links them.
//
// Mirrors oxmpl_b200/csrc/Makefile: `-fmad=false` is mandatory (the reference never fuses
// a*b+c, and neither may the kernels), `-ffp-contract=off` covers the host-side threshold code.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    // the CUDA sources live in the liboxg tree; OXG_CSRC overrides the default checkout layout
    let csrc = env::var("OXG_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| root.join("../../oxmpl_b200/csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    // every translation unit of oxmpl_b200/csrc/Makefile (tests/test_abi.py keeps the two lists equal)
    let sources = [
        "oxg_api.cu", "oxg_nn.cu", "oxg_validity.cu", "oxg_prm.cu", "oxg_planner.cu", "oxg_peak.cu",
        "oxg_so3.cu", "oxg_tree.cu", "oxg_dist.cu", "oxg_query.cu",
    ];
    for hdr in ["oxg_internal.cuh", "oxg_math.cuh", "oxg_tma.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(hdr).display());
    }
    for hdr in ["oxg.h", "oxg_so3.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join("../../include").join(hdr).display());
    }
    let mut objects = Vec::new();
    for src in sources {
        let obj = out.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            // one plain nvcc invocation per file: --split-compile makes ptxas's register allocation
            // of the scan kernels vary from build to build
            .args(["-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off"])
            .arg("-c")
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: set NVCC");
        assert!(status.success(), "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objects.push(obj);
    }
    let lib = out.join("liboxg.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=oxg");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl"); // oxg_dist.cu resolves NCCL with dlopen at run time
}
