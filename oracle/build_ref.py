#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- compiles the UNMODIFIED reference extensions from the sources where they lie under
/root/reference into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).  Nothing is copied into the repo.

    python oracle/build_ref.py [cpu] [binary_cuda] [q_linear_cuda] [functions_cuda]

  binary_linear_cpp   (CPU, OpenMP)  bitorch_engine/layers/qlinear/binary/cpp/binary_linear.cpp
                      -> bit-exact oracle + CPU baseline for the binary path
  binary_linear_cuda  (sm_100a)      bitorch_engine/layers/qlinear/binary/cuda/{binary_linear_cuda.cpp,..._kernel.cu}
  q_linear_cuda       (sm_100a)      bitorch_engine/layers/qlinear/nbit/cuda/{q_linear_cuda.cpp, mpq_..., mbwq_...}.cu
                      -> the reference CUDA path, run on the GPU box as the secondary oracle and as the kernel-to-beat
  functions_cuda      (sm_100a)      bitorch_engine/functions/cuda/{functions_cuda.cpp, functions_cuda_kernel.cu}
                      -> bit-exact oracle of the q4 / sign-bit wire formats on the GPU box
The reference's own build helper is bypassed on purpose (it appends -ccbin=/usr/bin/gcc-11, which does not exist here;
bitorch_engine/utils/cuda_extension.py:94-97)."""
import os
import sys

REF = "/root/reference/bitorch_engine"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _load(name, sources, cuda=False, extra_include=()):
    from torch.utils.cpp_extension import load
    os.environ.setdefault("CXX", "/usr/bin/g++")
    os.environ.setdefault("CC", "/usr/bin/gcc")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    build_dir = os.path.join(OUT, name)
    os.makedirs(build_dir, exist_ok=True)
    kw = dict(name=name + "_ref", sources=sources, build_directory=build_dir, verbose=False,
              extra_cflags=["-O3", "-fopenmp", "-Wno-deprecated-declarations"],
              extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"],
              extra_include_paths=list(extra_include))
    if cuda:
        kw["extra_cuda_cflags"] = ["-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-DARCH_SM_80",
                                   "-Xcompiler", "-fopenmp", "-Wno-deprecated-declarations",
                                   "-ccbin", "/usr/bin/g++"]
    return load(**kw)


def build_ref(what=("cpu",)):
    if not os.path.isdir(REF):
        return {}
    mods = {}
    if "cpu" in what:
        mods["binary_linear_cpp"] = _load("binary_linear_cpp",
                                          [f"{REF}/layers/qlinear/binary/cpp/binary_linear.cpp"])
    if "binary_cuda" in what:
        d = f"{REF}/layers/qlinear/binary/cuda"
        mods["binary_linear_cuda"] = _load("binary_linear_cuda",
                                           [f"{d}/binary_linear_cuda.cpp", f"{d}/binary_linear_cuda_kernel.cu"], cuda=True)
    if "q_linear_cuda" in what:
        d = f"{REF}/layers/qlinear/nbit/cuda"
        mods["q_linear_cuda"] = _load("q_linear_cuda",
                                      [f"{d}/q_linear_cuda.cpp", f"{d}/mpq_linear_cuda_kernel.cu",
                                       f"{d}/mbwq_linear_cuda_kernel.cu"], cuda=True, extra_include=[f"{d}/exl2"])
    if "functions_cuda" in what:
        d = f"{REF}/functions/cuda"
        mods["functions_cuda"] = _load("functions_cuda", [f"{d}/functions_cuda.cpp", f"{d}/functions_cuda_kernel.cu"],
                                       cuda=True)
    return mods


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref (no compilation; works on the GPU box)."""
    import glob
    import importlib.util
    import torch  # noqa: F401  (the .so links against libtorch)
    hits = glob.glob(os.path.join(OUT, name, f"{name}_ref*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location(f"{name}_ref", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    what = tuple(sys.argv[1:]) or ("cpu",)
    for k, v in build_ref(what).items():
        print("built", k, [a for a in dir(v) if not a.startswith("_")])
