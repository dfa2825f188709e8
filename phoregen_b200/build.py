"""Build libphoregen_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m phoregen_b200.build            # incremental
    python -m phoregen_b200.build --force
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphoregen_b200.so")
SOURCES = ["pg_graph.cu", "pg_gemm.cu", "pg_gemm_tc.cu", "pg_attn.cu", "pg_trip_tc.cu", "pg_bond_tc.cu", "pg_knn_tc.cu", "pg_model.cu", "pg_transition.cu"]
EXTRA = os.environ.get("PG_NVCC_EXTRA", "").split()     # e.g. PG_NVCC_EXTRA=-DPG_TRIP_TRACE for the phase tracer
NVCC_FLAGS = [*EXTRA, 
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "phoregen_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    stamp_file = os.path.join(CSRC, ".build_stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(log)[-6000:])
    if verbose:
        print("\n".join(log))
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB, *objs, "-lcudart"])
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
