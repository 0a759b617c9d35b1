"""In-tree build of libb200ocr.so (nvcc, sm_100a only).  `python -m onnxocr_b200.build`."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200ocr.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off",
          "-Xcudafe", "--diag_suppress=177"]
# (source, extra flags).  Geometry / integer post-processing must not contract a*b+c into FMA:
# its float arithmetic mirrors OpenCV's and Clipper's scalar x86 code operation by operation.
SOURCES = [
    ("onnx_reader.cc", []),
    ("graph.cc", []),
    ("engine.cu", []),
    ("stages.cu", ["-fmad=false"]),
    ("dbpost.cu", ["-fmad=false"]),
    ("abi.cu", []),
]


def _digest(path, flags):
    h = hashlib.sha1(" ".join(flags).encode())
    for f in sorted(os.listdir(CSRC)) + ["../../include/b200ocr.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src, extra, verbose):
    obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
    flags = ARCH + COMMON + extra
    stamp = obj + ".sha1"
    dig = _digest(src, flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = ["nvcc"] + flags + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose and r.stderr.strip():
        print(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj


def build(verbose=False, force=False):
    """Compiles every TU for sm_100a and links libb200ocr.so; a no-op when sources are unchanged."""
    stamp = LIB + ".stamp"
    total = _digest("all", ARCH + COMMON + [repr(SOURCES)])
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == total:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.unlink(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(lambda s: _compile(s[0], s[1], verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xlinker", "--exclude-libs,ALL"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(total)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
