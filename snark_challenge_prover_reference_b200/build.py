"""In-tree build of the CUDA library (sm_100a only): csrc/*.cu -> libb200groth16.so next to this file.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo snapshot to the GPU box.
    python -m snark_challenge_prover_reference_b200.build [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
# B200_VARIANT=<name>: build into build_<name>/ and libb200groth16_<name>.so (kernel variants selected with
# B200_NVCC_EXTRA, loaded with B200_LIB=<that .so>); the default build is untouched
_VAR = os.environ.get("B200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "libb200groth16%s.so" % ("_" + _VAR if _VAR else ""))
SOURCES = ["msm_g_mnt6g2.cu", "devops_g_mnt6g2.cu", "msm_g_mnt4g2.cu", "devops_g_mnt4g2.cu", "msm_g_mnt4g1.cu", "msm_g_mnt6g1.cu",
           "devops_g_mnt4g1.cu", "devops_g_mnt6g1.cu", "devops.cu", "msm.cu", "ntt.cu", "capi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-mbmi2", "-Xcompiler", "-madx", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
         "-ccbin", "/usr/bin/g++"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "b200_groth16.h"))
    return hs


def _flags():
    return FLAGS + os.environ.get("B200_NVCC_EXTRA", "").split()


def _stamp_matches():
    """build/flags.stamp records the effective nvcc flag list (B200_NVCC_EXTRA included): objects compiled with
    another flag set are stale even when their mtimes are newer than the sources."""
    want = " ".join(_flags())
    path = os.path.join(OBJ, "flags.stamp")
    have = open(path).read() if os.path.exists(path) else None
    if have != want:
        for f in os.listdir(OBJ):
            if f.endswith(".o"):
                os.remove(os.path.join(OBJ, f))
        with open(path, "w") as fh:
            fh.write(want)
        return False
    return True


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if _newer(obj, [path] + _headers()):
        return obj, ""
    cmd = [NVCC] + _flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout[-4000:], r.stderr[-8000:]))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
        if os.path.exists(LIB):
            os.remove(LIB)
    _stamp_matches()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    logs = "\n".join(l for _, l in results if l)
    if not _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lpthread",
                                                     "-ccbin", "/usr/bin/g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB, logs


def build_driver():
    """Host-side C++ (the reference is C++): the B:: bundle over the C ABI + the command-line driver ->
    bin/cuda_prover_piecewise, linked against the in-tree CUDA library (rpath $ORIGIN/..)."""
    host = os.path.join(CSRC, "host")
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    exe = os.path.join(bindir, "cuda_prover_piecewise")
    srcs = [os.path.join(host, "b200_bundle.cpp"), os.path.join(host, "prover_main.cpp")]
    deps = srcs + [os.path.join(host, "b200_bundle.hpp"), os.path.join(host, "prover_reference_functions.hpp"), LIB]
    if _newer(exe, deps):
        return exe
    cmd = ["/usr/bin/g++", "-std=c++14", "-O2", "-I", host, "-I", os.path.join(ROOT, "include")] + srcs + [
        "-L", HERE, "-lb200groth16", "-Wl,-rpath,$ORIGIN/..", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host driver build failed:\n" + r.stderr[-4000:])
    return exe


if __name__ == "__main__":
    lib, logs = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if logs:
        print(logs)
    print("built", lib)
    print("built", build_driver())
