"""Developer tool: builds a kernel-variant copy of the library (extra -D flags) next to the product one, for A/B runs with
BP_B200_LIB=<path> (api.py honours it).  python tools/build_variant.py NAME -DSB_SEG=256 ...  ->  _obj/variant_NAME/libbp_b200.so"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.join(HERE, ".."))
from bulletproofs_r1cs_gadgets_b200 import build as B
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(B.OBJ, "variant_" + name); os.makedirs(out, exist_ok=True)
def cc(u):
    subprocess.run([B.NVCC] + [f for f in B.FLAGS if f not in ("-Xptxas", "-v")] + extra + ["-c", os.path.join(B.CSRC, u + ".cu"), "-o", os.path.join(out, u + ".o")], check=True)
with ThreadPoolExecutor(8) as ex: list(ex.map(cc, B.UNITS))
so = os.path.join(out, "libbp_b200.so")
subprocess.check_call([B.NVCC, "-shared", "-o", so] + [os.path.join(out, u + ".o") for u in B.UNITS] + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(so)
