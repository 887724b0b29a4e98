"""Builds libbp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bulletproofs_r1cs_gadgets_b200.build [--force]

One object per translation unit, compiled in parallel.  Whether the objects are current is decided by CONTENT: a SHA-256 over
every source and header (and the compiler flags) is kept in _obj/build_stamp.json next to the objects; any difference -- or a
missing object, stamp or library -- rebuilds everything.  (Binaries are not tracked by git; a tree copied with its _obj/ and
.so keeps them only if they were built from exactly these sources.)  The stamp also records nvcc's wall time.
"""
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
SO = os.path.join(PKG, "libbp_b200.so")
UNITS = ["engine", "gadgets", "capi", "tree", "k_msm", "k_fold", "k_points", "k_transcript", "k_scalar", "k_table", "k_sorted"]
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


STAMP = os.path.join(OBJ, "build_stamp.json")


def source_digest():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cu")))
    files.append(os.path.join(PKG, "..", "include", "bp_b200.h"))
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0" + open(f, "rb").read() + b"\0")
    return h.hexdigest()


def _compile(unit):
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stderr[-4000:]))
    with open(os.path.join(OBJ, unit + ".ptxas.txt"), "w") as f:
        f.write(r.stderr)
    return unit, True, r.stderr


def is_current():
    try:
        stamp = json.load(open(STAMP))
    except (OSError, ValueError):
        return False
    return stamp.get("digest") == source_digest() and os.path.exists(SO) and all(os.path.exists(os.path.join(OBJ, u + ".o")) for u in UNITS)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if not force and is_current():
        if verbose:
            print("up to date (sources match %s)" % STAMP)
        return SO
    if os.path.exists(STAMP):
        os.remove(STAMP)
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, UNITS))
    objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
    subprocess.check_call([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    wall = time.time() - t0
    json.dump({"digest": source_digest(), "nvcc_wall_seconds": round(wall, 1), "units": UNITS, "flags": FLAGS,
               "nvcc": subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]}, open(STAMP, "w"), indent=1)
    if verbose:
        print("compiled %d units and linked in %.1f s" % (len(results), wall))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
