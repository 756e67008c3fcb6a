"""Builds the C-ABI shared library in-tree with nvcc for sm_100a (no torch headers involved).

    python -m simple_pose_b200.build        # or __graft_entry__.build()

The .so lands in ``simple_pose_b200/lib/`` (git-ignored, but it travels to the GPU box with the
working tree). nvcc cross-compiles without a GPU, so this runs in the CPU-only build container.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIBNAME = "libsimple_pose_b200.so"
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _fingerprint():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(os.path.dirname(PKG), "include", "simple_pose_b200.h")]
    root = os.path.dirname(PKG)
    for f in files:
        with open(f, "rb") as fh:       # path relative to the checkout: the same sources hash the same wherever the tree lives
            h.update(os.path.relpath(f, root).encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the C-ABI library cannot be built")


def build(force=False, verbose=False):
    """Compiles every csrc/*.cu to an object file (in parallel) and links them into the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, LIBNAME + ".sha256")
    fp = _fingerprint()
    if not force and os.path.isfile(lib_path()) and os.path.isfile(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == fp:
                return lib_path()
    nvcc = find_nvcc()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
        return obj, proc.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _sources()))
    if verbose:
        for _, out in results:
            print(out)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", lib_path()] + [obj for obj, _ in results]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
