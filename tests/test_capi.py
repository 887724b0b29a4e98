"""The shipped C-ABI library: loads, exports every entry point include/bp_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("bp_gens_new", "bp_prover_new", "bp_prover_commit", "bp_verifier_commit", "bp_cs_multiply", "bp_cs_allocate_multiplier",
                 "bp_cs_allocate_single", "bp_cs_evaluate_lc", "bp_cs_constrain", "bp_cs_num_constraints", "bp_cs_num_multipliers",
                 "bp_prover_prove", "bp_verifier_verify", "bp_circuit_compile", "bp_prove_batch", "bp_verify_batch", "bp_verify_batch_combined", "bp_verify_batch_combined_device", "bp_proof_to_wire", "bp_proof_from_wire", "bp_gadget_vsmt4_verif", "bp_gadget_poseidon_hash_4", "bp_poseidon_hash_4", "bp_prove_batch_device", "bp_prove_stream_begin", "bp_prove_stream_finish", "bp_prove_stream_begin_host", "bp_prove_stream_finish_host", "bp_vsmt2_new", "bp_vsmt2_free", "bp_vsmt2_depth", "bp_vsmt2_num_nodes", "bp_vsmt2_root", "bp_vsmt2_empty_hashes", "bp_vsmt2_update_batch", "bp_vsmt2_get_batch", "bp_vsmt2_witness_batch", "bp_vsmt2_witness_batch_device", "bp_poseidon_hash_2_batch",
                 "bp_msm_gens_device"):
        assert must in names


def test_library_exports_every_declared_symbol(product_so):
    lib = C.CDLL(product_so)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.bp_version() >= 1


def test_python_layer_declares_every_prototype(product_so):
    """the ctypes layer takes argtypes/restype from the header: every declared function is covered, arity is enforced"""
    from bulletproofs_r1cs_gadgets_b200 import api
    protos = api.header_prototypes()
    assert sorted(n for n, _, _ in protos) == declared_functions()
    lib = C.CDLL(product_so)
    api._declare_prototypes(lib)
    for name, ret, params in protos:
        fn = getattr(lib, name)
        assert fn.argtypes is not None and len(fn.argtypes) == len(params), name
    assert lib.bp_launch_count.restype is C.c_int64 and lib.bp_vsmt2_free.restype is None
    with pytest.raises((TypeError, C.ArgumentError)):
        lib.bp_gens_new(16)  # wrong number of arguments is refused instead of reading a garbage pointer


def test_emulation_build_is_not_the_product(product_so):
    """the product library is a CUDA binary for sm_100a (no host-emulation launcher inside)"""
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % product_so).read()
    assert "sm_100a" in out
    api_src = open(os.path.join(ROOT, "bulletproofs_r1cs_gadgets_b200", "api.py")).read()
    assert "emul" not in api_src.replace("emulation build", "")  # the Python layer never reaches for the emulation library


def test_fails_loudly_without_gpu(product_so):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = C.CDLL(product_so)
    h = C.c_void_p()
    assert lib.bp_gens_new(C.c_uint32(16), C.byref(h)) == 7  # BP_ERR_NO_DEVICE: no CPU fallback
    out = (C.c_uint8 * 32)()
    assert lib.bp_selftest_device(1, (C.c_uint8 * 32)(), C.c_size_t(32), out, C.c_size_t(32)) == 7


def test_python_layer_requires_the_library(tmp_path, monkeypatch):
    from bulletproofs_r1cs_gadgets_b200 import api
    saved = api._lib
    api._lib = None
    monkeypatch.setattr(api, "DEFAULT_SO", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError):
        api.load()
    api._lib = saved


def _build_and_run_smoke(tmp_path, so):
    """tests/capi_smoke.c: a C11 translation unit that includes include/bp_b200.h, compiled with -Wall -Wextra -Werror and linked
    against `so` -- the header's prototypes and the library's entry points have to agree for this to compile, link and pass"""
    exe = str(tmp_path / "capi_smoke")
    d, name = os.path.dirname(so), os.path.basename(so)
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "capi_smoke.c"), "-o", exe,
           "-L" + d, "-l:" + name, "-Wl,-rpath," + d]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True)


def test_c_translation_unit_against_the_kernel_bodies(tmp_path, emul_so):
    r = _build_and_run_smoke(tmp_path, emul_so)
    assert r.returncode == 0 and "capi_smoke ok" in r.stdout, r.stdout + r.stderr


def test_c_translation_unit_links_against_the_product(tmp_path, product_so):
    """on the CPU box the product library links and runs up to its first device call, which must report BP_ERR_NO_DEVICE (7)"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: tests/test_gpu.py runs the full program")
    except ImportError:
        pass
    r = _build_and_run_smoke(tmp_path, product_so)
    assert r.returncode == 1 and "bp_gens_new(16, &g) -> 7" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_translation_unit_on_the_device(tmp_path, product_so):
    r = _build_and_run_smoke(tmp_path, product_so)
    assert r.returncode == 0 and "capi_smoke ok" in r.stdout, r.stdout + r.stderr
