// build.rs -- compiles the sm_100a kernels with nvcc and links them into the Rust binary.
//
// Inputs : ../../csrc/*.cu (plan, esc, esc_cta_bitonic, fused, transpose, heavy, heavy_smem, engine)
//          + ../../csrc/{common,sort,cta_common}.cuh + ../../../include/spada_b200.h
// Output : $OUT_DIR/libspada_b200.a, linked statically together with cudart.
// There is exactly one code path: sm_100a.  No Triton, no multi-backend dispatch, no CPU fallback.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".to_string());
    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".to_string());

    let mut objects = vec![];
    for unit in ["plan", "esc", "esc_cta_bitonic", "fused", "transpose", "heavy", "heavy_smem", "engine"].iter() {
        let src = csrc.join(format!("{}.cu", unit));
        let obj = out.join(format!("{}.o", unit));
        println!("cargo:rerun-if-changed={}", src.display());
        let status = Command::new(&nvcc)
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
            .args(&["--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"])
            .arg("-c")
            .arg(&src)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc (set NVCC=/path/to/nvcc)");
        assert!(status.success(), "nvcc failed on {}", src.display());
        objects.push(obj);
    }
    println!("cargo:rerun-if-changed={}", csrc.join("common.cuh").display());
    println!("cargo:rerun-if-changed={}", manifest.join("../../../include/spada_b200.h").display());

    let lib = out.join("libspada_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().expect("ar");
    assert!(status.success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=spada_b200");
    println!("cargo:rustc-link-search=native={}", cuda_lib);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
