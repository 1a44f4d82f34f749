"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/plonk_b200.h
declares, and REFUSES to compute without a CUDA device (no CPU fallback).  No compute calls are made here
when a GPU is absent."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "plonk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(host):
    lib = host.lib()
    names = declared_symbols()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/plonk_b200.h but not exported: {missing}"
    lib.pb_abi_version.restype = C.c_int
    assert lib.pb_abi_version() == 2


def dropin_declared_functions():
    """Functions the ten drop-in headers declare (everything between their extern "C" guards)."""
    names = {}
    for h in ("poly", "matrix", "g1", "g2", "gt", "pairing", "srs", "constraints", "plonk"):
        text = open(os.path.join(ROOT, "include", h + ".h")).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        m = re.search(r'extern "C" \{\s*#endif(.*?)#ifdef __cplusplus\s*\}', text, flags=re.S)
        assert m, h
        for fn in re.findall(r"\b([a-z][a-z0-9_]*)\s*\([^;{]*\)\s*;", m.group(1)):
            names[fn] = h
    return names


def test_library_exports_every_dropin_function(host):
    """include/{poly,matrix,g1,g2,gt,pairing,srs,constraints,plonk}.h only DECLARE the reference's functions; the shared
    library must define every one of them (the reference defines them in its headers: src/*.h)."""
    lib = host.lib()
    names = dropin_declared_functions()
    assert len(names) >= 60, len(names)
    for must in ("poly_mul", "poly_divide", "matrix_inv", "g1_mul", "g2_add", "gtp_pow", "pairing", "srs_eval_at_s", "eval_expr",
                 "constraints_satisfy", "plonk_new", "interpolate_at_h", "plonk_prove", "poly_print", "copy_constraints_to_roots"):
        assert must in names, must
    missing = [f"{n} ({h}.h)" for n, h in names.items() if not hasattr(lib, n)]
    assert not missing, missing
    assert not hasattr(lib, "matrix_equal")      # the reference's tests define their own (plonk-test.c:11, matrix-test.c:4)


def test_every_dev_entry_point_has_a_host_twin():
    names = set(declared_symbols())
    for n in names:
        if n.endswith("_dev") and n not in ("pb_tally_dev", "pb_plonk_verify_completed_dev", "pb_peak_probe_dev", "pb_plonk_prove_verify_ex_dev", "pb_config2_items_dev",
                                             "pb_gather_completed_dev", "pb_synth_batch_dev", "pb_plonk_prove_verify_tally_dev"):
            assert n[:-4] in names, n


def test_no_cpu_fallback(host, W):
    """On a box without a GPU every compute entry point must fail loudly with PB_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on CPU-only hosts")
    a = np.arange(16, dtype=np.uint8)
    with pytest.raises(host.PlonkB200Error) as e:
        host.hf_add(a, a)
    assert e.value.code == host.PB_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(host.PlonkB200Error) as e:
        host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    assert e.value.code == host.PB_ERR_NO_DEVICE
    with pytest.raises(host.PlonkB200Error):
        host.pairing(np.array([[1, 2, 0]], np.uint8), np.array([[36, 31]], np.uint8))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under plonk.c_b200/ or include/ may reference it."""
    bad = []
    for base in ("plonk.c_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    text = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\".*oracle/|libplonk_port|libref_oracle", text, flags=re.M):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def test_workload_generator_is_counter_based(W):
    a = W.make_batch(9, 100, 50)
    b = W.make_batch(9, 0, 150)
    for x, y in zip(a, b):
        assert np.array_equal(x, y[100:150])
    t = W.satisfying_witnesses()
    assert t.shape == (289, 12) and list(t[71][:3]) == [3, 4, 5]
    x, y, z = t[:, 0].astype(int), t[:, 1].astype(int), t[:, 2].astype(int)
    assert np.all((x * x + y * y - z * z) % 17 == 0)
    wit, rnd, chal, u = W.make_batch(1, 0, 1000, "NZ")
    assert rnd.min() >= 1 and chal.min() >= 1 and rnd.max() <= 16
