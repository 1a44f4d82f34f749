"""Circuit front-end, end to end: expression -> eval_expr -> GATE_LIST (the reference's API,
src/constraints.h:227-309) -> pb_circuit_from_gates -> 44-byte circuit -> context -> proofs, against the compiled reference
given the same circuit bytes.  The lowering is host code and is tested here without a GPU; proving needs one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity_suite as ps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "plonk.c_b200")


def _frontend(host, tmp_path):
    """Compile and run tests/c/circuit_frontend.c (plain C against include/*.h); parse its circuits."""
    exe = str(tmp_path / "circuit_frontend")
    with open(os.path.join(ROOT, "tests", "c", "circuit_frontend.c")) as f:
        r = subprocess.run(["gcc", "-I", INC, "-x", "c", "-o", exe, "-", "-L", LIBDIR, "-lplonk_b200", f"-Wl,-rpath,{LIBDIR}"],
                           stdin=f, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, run.stdout + run.stderr
    out = {}
    for line in run.stdout.splitlines():
        t = line.split()
        if t[0] == "five_gates":
            out["five_gates"] = dict(gates=int(t[2]), rc=int(t[4]))
            continue
        g = int(t[2])
        ia, ib, ic = t.index("a"), t.index("b"), t.index("c")
        out[t[0]] = dict(a=[int(v) for v in t[ia + 1:ia + 1 + g]], b=[int(v) for v in t[ib + 1:ib + 1 + g]],
                         c=[int(v) for v in t[ic + 1:ic + 1 + g]], n_vars=int(t[t.index("vars") + 1]),
                         circuit=np.array([int(v) for v in t[t.index("circuit") + 1:]], np.uint8))
    return out


def _witness(host, c, values):
    idx = [np.array(c[k], np.uint64) for k in "abc"]
    n = values.shape[0]
    wit = np.zeros((n, 12), np.uint8)
    vals = np.ascontiguousarray(values, np.uint8)
    host._check(host.lib().pb_witness_from_values(*(i.ctypes.data_as(C.c_void_p) for i in idx), C.c_size_t(len(c["a"])),
                                                  vals.ctypes.data_as(C.c_void_p), C.c_size_t(vals.shape[1]), wit.ctypes.data_as(C.c_void_p),
                                                  C.c_size_t(n)))
    return wit


def test_lowering_known_answers(host, W, tmp_path):
    c = _frontend(host, tmp_path)
    # gate by gate in the order of plonk-test.c: byte for byte the circuit the reference's test writes by hand (plonk-test.c:157-213)
    assert np.array_equal(c["plonk_test"]["circuit"], W.PLONK_TEST_CIRCUIT)
    assert c["plonk_test"]["a"] == [0, 2, 4, 1] and c["plonk_test"]["b"] == [0, 2, 4, 3] and c["plonk_test"]["c"] == [1, 3, 5, 5]
    # one expression tree: mul(x,x) mul(y,y) sum mul(z,z); the sum's output and z*z's output share a copy cycle
    e = c["pythagoras_expr"]
    assert e["a"] == [0, 2, 1, 5] and e["b"] == [0, 2, 3, 5] and e["c"] == [1, 3, 4, 6] and e["n_vars"] == 7
    circ = e["circuit"]
    assert circ[:20].reshape(5, 4).tolist() == [[0, 0, 1, 0], [0, 0, 1, 0], [16, 16, 16, 16], [1, 1, 0, 1], [0, 0, 0, 0]]
    typ, idx = circ[20:].reshape(3, 8)[:, :4], circ[20:].reshape(3, 8)[:, 4:]
    sigma = {(t, i): (int(typ[t, i]), int(idx[t, i]) - 1) for t in range(3) for i in range(4)}
    assert sorted(sigma.values()) == sorted(sigma.keys())                                 # a permutation of the 12 positions
    var_at = {(0, i): e["a"][i] for i in range(4)} | {(1, i): e["b"][i] for i in range(4)} | {(2, i): e["c"][i] for i in range(4)}
    canon = lambda v: 4 if v == 6 else v                                                # variables 4 and 6 were asserted equal
    assert all(canon(var_at[p]) == canon(var_at[q]) for p, q in sigma.items())          # it only ever links equal variables
    for v in set(map(canon, var_at.values())):                                          # and each variable's positions form ONE cycle
        pos = [p for p in var_at if canon(var_at[p]) == v]
        seen, p = set(), pos[0]
        while p not in seen:
            seen.add(p)
            p = sigma[p]
        assert seen == set(pos)
    assert c["five_gates"]["gates"] == 5 and c["five_gates"]["rc"] == host.PB_ERR_ARG


def test_lowered_circuit_agrees_with_the_reference_on_cpu(host, ref, W, tmp_path):
    """Given the lowered circuit bytes, the unmodified reference proves satisfying witnesses and aborts on the others: the
    lowering produces a circuit the reference itself accepts."""
    e = _frontend(host, tmp_path)["pythagoras_expr"]
    tri = W.satisfying_witnesses()[:, :3].astype(np.int64)                                # (x, y, z) with x^2 + y^2 = z^2
    x, y, z = tri[:, 0], tri[:, 1], tri[:, 2]
    vals = np.stack([x, x * x % 17, y, y * y % 17, (x * x + y * y) % 17, z, z * z % 17], axis=1).astype(np.uint8)
    wit = _witness(host, e, vals)
    n = wit.shape[0]
    g1s, g2 = W.generator_srs(9)
    _, rnd, chal, _ = W.make_batch(4, 0, n, "NZ")
    _, status = ref.plonk_prove_batch(e["circuit"], g1s, g2, wit, rnd, chal)
    assert (status != 1).all() and (status == 0).sum() > n // 3
    bad = wit.copy()
    bad[:, 10] = (bad[:, 10] + 1) % 17                     # break the sum gate's output
    _, status = ref.plonk_prove_batch(e["circuit"], g1s, g2, bad, rnd, chal)
    assert (status == 1).all()


@pytest.mark.gpu
def test_expression_to_proofs_on_gpu(host, oracle, W, tmp_path):
    """x*x + y*y = z*z authored with eval_expr, lowered, proved and verified on the GPU: 10 000 witnesses (satisfying
    triples x random blinding / challenges, plus unsatisfying ones), every byte against the oracle on the same circuit."""
    circuits = _frontend(host, tmp_path)
    n = 10000
    tri = W.satisfying_witnesses()[:, :3].astype(np.int64)
    rng = np.random.default_rng(12)
    pick = tri[rng.integers(0, len(tri), n)]
    x, y, z = pick[:, 0], pick[:, 1], pick[:, 2]
    g1s, g2 = W.generator_srs(9)
    for name, vals in (("plonk_test", np.stack([x, x * x % 17, y, y * y % 17, z, z * z % 17], axis=1)),
                       ("pythagoras_expr", np.stack([x, x * x % 17, y, y * y % 17, (x * x + y * y) % 17, z, z * z % 17], axis=1))):
        c = circuits[name]
        wit = _witness(host, c, vals.astype(np.uint8))
        wit[::9, 8] = (wit[::9, 8] + 1) % 17                # every ninth witness violates gate 0
        _, rnd, chal, u = W.make_batch(13, 0, n, "U17")
        pk = host.Plonk(c["circuit"], g1s, g2)
        got = pk.prove_verify(wit, rnd, chal, u)
        rp, rs = oracle.plonk_prove_batch(c["circuit"], g1s, g2, wit, rnd, chal)
        rv, _ = oracle.plonk_verify_batch(c["circuit"], g1s, g2, rp, chal, u)
        rv = np.where(rs == 0, rv, 0xFF).astype(np.uint8)
        ps.eq(name, got, (rp, rs, rv))
        assert (rs[::9] == 1).all() and (rs == 0).sum() > n // 3
