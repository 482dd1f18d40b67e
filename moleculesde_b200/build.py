"""In-tree build of `libmolsde_b200.so` with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmolsde_b200.so")
FASTCALL_PATH = os.path.join(_HERE, "_molsde_fastcall.so")   # CPython extension: low-overhead call path (csrc/fastcall.c)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(os.path.dirname(_HERE), "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_fastcall(force: bool = False) -> str:
    """gcc-compiled CPython extension next to the library (x86-64 SysV); rebuilt when its source is newer."""
    import sysconfig
    src = os.path.join(CSRC, "fastcall.c")
    if not force and os.path.exists(FASTCALL_PATH) and os.path.getmtime(FASTCALL_PATH) >= os.path.getmtime(src):
        return FASTCALL_PATH
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], src, "-o", FASTCALL_PATH]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed on fastcall.c:\n{r.stdout}")
    return FASTCALL_PATH


def build(force: bool = False, verbose: bool = False) -> str:
    build_fastcall(force)
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(_HERE, "_build", os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        if os.environ.get("MOLSDE_PROF") == "1":
            cmd += ["-DMOLSDE_PROF"]
        if os.environ.get("MOLSDE_CFLAGS"):   # A/B switches of single kernels (e.g. -DMOLSDE_SINCOS_REF_ROUNDING)
            cmd += os.environ["MOLSDE_CFLAGS"].split()
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-o", LIB_PATH] + objs  # static cudart (nvcc default): independent of torch's runtime
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
