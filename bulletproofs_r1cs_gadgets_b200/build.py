"""Builds libbp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bulletproofs_r1cs_gadgets_b200.build [--force]

One object per translation unit, compiled in parallel; relinked only when an object changed.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
SO = os.path.join(PKG, "libbp_b200.so")
UNITS = ["engine", "gadgets", "capi", "tree", "k_msm", "k_fold", "k_points", "k_transcript", "k_scalar", "k_table", "k_sorted"]
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    hs.append(os.path.join(PKG, "..", "include", "bp_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(unit, force, hdr_time):
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
        return unit, False, ""
    r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stderr[-4000:]))
    with open(os.path.join(OBJ, unit + ".ptxas.txt"), "w") as f:
        f.write(r.stderr)
    return unit, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest_header()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda u: _compile(u, force, hdr_time), UNITS))
    changed = any(c for _, c, _ in results)
    if changed or not os.path.exists(SO):
        objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
        subprocess.check_call([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    if verbose:
        for u, c, log in results:
            print(u, "compiled" if c else "up to date")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
