"""Builds tests/host/libb200geom_host.so: the product's dbgeom.h compiled for the CPU (g++), so the geometry
that runs inside the CUDA kernels can be fuzzed against cv2 without a GPU.  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libb200geom_host.so")


def build():
    src = os.path.join(HERE, "geom_host.cc")
    hdr = os.path.join(HERE, "..", "..", "onnxocr_b200", "csrc", "dbgeom.h")
    if os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, src]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build())
