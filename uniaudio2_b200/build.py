"""In-tree build of the sm_100a C-ABI library (no JIT cache: the .so must travel with the repo snapshot).

    python -m uniaudio2_b200.build            # nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...

Each .cu is compiled to an object in parallel, then linked into uniaudio2_b200/libua2_b200.so.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libua2_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


EXTRA = {}  # per-file extra flags (none: every kernel, the tcgen05 mainloop included, is this repo's own source)


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, hdig, verbose):
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".stamp"
    dig = hashlib.sha1(open(src, "rb").read() + hdig.encode() + " ".join(EXTRA.get(os.path.basename(src), [])).encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + EXTRA.get(os.path.basename(src), []) + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    open(stamp, "w").write(dig)
    return obj, True


def stale_sources():
    """Sources whose object file is missing or was compiled from different text / headers / flags than the tree holds now, plus
    the library itself when it is older than an object - i.e. what `build()` would redo.  Empty list = the .so is current."""
    hdig = _headers_digest()
    stale = []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        dig = hashlib.sha1(open(src, "rb").read() + hdig.encode() + " ".join(EXTRA.get(os.path.basename(src), [])).encode()).hexdigest()
        if not (os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == dig):
            stale.append(os.path.basename(src))
        elif not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(obj):
            stale.append(os.path.basename(LIB))
    return sorted(set(stale))


def build(force: bool = False, verbose: bool = False) -> str:
    if force and os.path.isdir(OBJ):
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    hdig = _headers_digest()
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, hdig, verbose), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
