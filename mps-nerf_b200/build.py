"""Build libmpsnerf_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpsnerf_b200.so")
SOURCES = ["abi.cu", "raygen.cu", "frame_prep.cu", "grid_build.cu", "sample_knn.cu", "deform.cu", "gather.cu", "composite.cu", "occupancy.cu",
           "dense_fp32.cu", "train_bwd.cu", "render_rays.cu", "dense_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


HASH_FILE = LIB + ".srchash"


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def _headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
        [os.path.join(ROOT, "include", "mpsnerf.h")]


def source_hash():
    """sha256 over every source, header and the compiler flags: what the shipped .so must have been built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS[:-4]).encode())
    for f in [os.path.join(CSRC, s) for s in SOURCES] + _headers():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_current():
    """True if the in-tree .so was built from the sources as they are now (mtime-independent: the .so travels to
    the GPU box in a snapshot whose timestamps mean nothing)."""
    try:
        with open(HASH_FILE) as fh:
            return os.path.exists(LIB) and fh.read().strip() == source_hash()
    except OSError:
        return False


def build(force=False, verbose=False):
    """Compile every .cu that is newer than its object, then link.  Returns the .so path.

    Safe under concurrent callers (torchrun ranks importing at once): the whole build runs under an exclusive file
    lock, late arrivals find the library current and return, and the .so is linked to a temporary name and moved
    into place atomically, so nobody can dlopen a half-written file."""
    import fcntl
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB
            return _build_locked(force, verbose, objdir)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose, objdir):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = _headers()
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB + ".tmp.%d" % os.getpid()
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-lcudart", "-ldl"])
    os.replace(tmp, LIB)
    with open(HASH_FILE + ".tmp", "w") as fh:
        fh.write(source_hash())
    os.replace(HASH_FILE + ".tmp", HASH_FILE)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
