"""Build recipes for the native libraries (in-tree, so the built .so files travel with the repo snapshot).

  csrc/liben234gpu.so   CUDA kernels + C ABI (include/en234gpu.h), sm_100a only
  csrc/liben234host.so  C++ host driver (include/en234host.h): .in parser, loads, static step loop
  csrc/en234_run        command-line driver:  en234_run <file.in> [options]
  oracle/liben234oracle.so  CPU oracle (test infrastructure only; built here, never linked into the product)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "en234_fea_b200", "csrc")
HOST = os.path.join(CSRC, "host")
ORACLE = os.path.join(ROOT, "oracle")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]
HOST_SRCS = ["read_input_file.cpp", "user_mesh.cpp", "loads.cpp", "staticstep.cpp", "host_capi.cpp", "user_element.cpp", "checkstiffness.cpp", "print_state.cpp", "explicit_dynamic_step.cpp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_gpu(force=False):
    out = os.path.join(CSRC, "liben234gpu.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "en234gpu.h"))
    if force or _newer(out, srcs):
        _run(["nvcc"] + NVCC_FLAGS + ["-o", out, "en234gpu.cu", "-ldl"], CSRC)
    return out


def build_host(force=False):
    out = os.path.join(CSRC, "liben234host.so")
    srcs = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [os.path.join(ROOT, "include", "en234host.h"),
                                                               os.path.join(CSRC, "liben234gpu.so")]
    cxx = os.environ.get("CXX", "g++")
    if force or _newer(out, srcs):
        _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out] +
             [os.path.join("host", s) for s in HOST_SRCS] +
             ["-L.", "-len234gpu", "-Wl,-rpath,$ORIGIN"], CSRC)
    exe = os.path.join(CSRC, "en234_run")
    if force or _newer(exe, srcs + [os.path.join(HOST, "main.cpp"), out]):
        _run([cxx, "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join("host", "main.cpp"),
              "-L.", "-len234host", "-len234gpu", "-Wl,-rpath,$ORIGIN"], CSRC)
    return out


def build_oracle(force=False):
    out = os.path.join(ORACLE, "liben234oracle.so")
    if force:
        _run(["make", "clean"], ORACLE)
    _run(["make", "-j8"], ORACLE)
    return out


def build_ctests(force=False):
    """tests/c/multi_check: plain-C driver of the multi-GPU entry points, linked against liben234gpu.so only."""
    src = os.path.join(ROOT, "tests", "c", "multi_check.c")
    out = os.path.join(ROOT, "tests", "c", "multi_check")
    if os.path.exists(src) and (force or _newer(out, [src, os.path.join(CSRC, "liben234gpu.so"), os.path.join(ROOT, "include", "en234gpu.h")])):
        _run([os.environ.get("CC", "gcc"), "-O2", "-Wall", "-o", out, src, "-I" + os.path.join(ROOT, "include"), "-L" + CSRC, "-len234gpu",
              "-lm", "-Wl,-rpath,$ORIGIN/../../en234_fea_b200/csrc"], ROOT)
    return out


def build_all(force=False):
    build_gpu(force)
    build_host(force)
    build_oracle(force)
    build_ctests(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built")
