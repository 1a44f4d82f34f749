"""CPU tests of the PRODUCT's kernel source: field.cuh / curve.cuh / prover.cuh / verifier.cuh are
__host__ __device__, so the exact code the GPU runs is compiled for the host (tests/hostcheck) and diffed
against the oracle here, where there is no GPU.  This is a debugging aid, not a product path."""
import pytest

import parity_suite as ps
from impls import HostcheckImpl


@pytest.fixture(scope="module")
def hc(oracle):
    return HostcheckImpl(oracle)


def test_hostcheck_known_answers(hc):
    ps.check_reference_known_answers(hc)


def test_hostcheck_groups(hc, oracle):
    ps.check_groups(hc, oracle)


def test_hostcheck_pairing(hc, oracle, W):
    ps.check_pairing(hc, oracle, W, n=30000)


def test_hostcheck_protocol(hc, oracle, W):
    ps.check_protocol(hc, oracle, W, n=20000)


def test_hostcheck_golden(hc, oracle, W):
    ps.check_golden_transcript(hc, W)
    ps.check_groups_golden(hc, W, oracle)
    ps.check_protocol_golden(hc, W)


def test_hostcheck_fast_paths(oracle, W):
    """The fast paths (pair-table commitments, fixed-base verifier tables, joint double-and-add) are only taken for a
    canonically encoded on-curve SRS; on that domain they must still be byte-identical to the oracle."""
    import util
    hcf = HostcheckImpl(oracle, fast=True)
    modes = [(m, f) for m, f in util.SRS_MODES.items()]
    ps.check_protocol(hcf, oracle, W, n=20000, modes=modes)
    ps.check_golden_transcript(hcf, W)


def test_hostcheck_shared_final_exponentiation(hc):
    """The verdict-only verifier compares two pairings with ONE final exponentiation (curve.cuh: pairings_equal17_c).
    The identity behind it is checked here on GT values directly: every f2 in GT against every 5th f1 (21 M pairs)."""
    import ctypes as C
    hc.lib.hc_check_shared_final_exp.restype = C.c_uint64
    assert hc.lib.hc_check_shared_final_exp(C.c_uint32(5)) == 0


def test_hostcheck_canonical_domain_variants(hc):
    """g1_add_c / g1_double_c / miller17_c (the fast verifier's leaner group law and Miller loop) against the any-input
    functions on their whole domain: all 102 x 102 pairs of canonical curve points, all 102 x 10201 (P, Q) pairs."""
    import ctypes as C
    import numpy as np
    pts = np.concatenate([ps.curve_points(), np.array([[0, 0, 1]], np.uint8)])
    assert pts.shape == (102, 3)
    hc.lib.hc_check_canonical_variants.restype = C.c_uint64
    assert hc.lib.hc_check_canonical_variants(pts.ctypes.data_as(C.c_void_p), C.c_uint32(len(pts))) == 0


def test_hostcheck_group_law_on_canonical_points(hc, oracle):
    """The exactness argument of the pair / wide / verifier tables and of the joint double-and-add: restricted to
    canonically encoded points of E(F_101), g1_add (which test_hostcheck_groups pins to the reference for any input) is
    commutative, associative, closed and has {0,0,1} as its neutral element -- as byte triples.  All 102^3 triples."""
    import ctypes as C
    import numpy as np
    pts = np.concatenate([ps.curve_points(), np.array([[0, 0, 1]], np.uint8)])
    hc.lib.hc_check_group_law.restype = C.c_uint64
    assert hc.lib.hc_check_group_law(pts.ctypes.data_as(C.c_void_p), C.c_uint32(len(pts))) == 0
    # and the function being exercised is the reference's, on exactly these inputs
    a = np.repeat(pts, len(pts), axis=0)
    b = np.tile(pts, (len(pts), 1))
    ps.eq("g1_add on all canonical pairs", hc.g1_op(0, a, b), oracle.g1_op(0, a, b))


def test_hostcheck_extreme_inputs_stay_within_the_reduction_bounds(oracle, W):
    """The test library is built with -DPB_CHECK_BOUNDS: every reduction and table look-up aborts the process if its
    input leaves the range the kernel relies on.  Inputs chosen to push the unreduced accumulators as high as they go
    (every byte 16, or 0, or alternating; random satisfying and unsatisfying witnesses; every SRS mode and table path) --
    the prover is straight-line, so even an unsatisfied witness runs through all five rounds."""
    import numpy as np
    import util
    assert HostcheckImpl(oracle).lib.hc_bounds_checked() == 1, "tests/hostcheck must be built with -DPB_CHECK_BOUNDS"
    C = W.PLONK_TEST_CIRCUIT
    rng = np.random.default_rng(99)
    n = 3000
    wit = rng.integers(0, 17, (n, 12), dtype=np.uint8)
    rnd = rng.integers(0, 17, (n, 9), dtype=np.uint8)
    chal = rng.integers(0, 17, (n, 5), dtype=np.uint8)
    for k, v in enumerate((16, 0, 1)):
        wit[k], rnd[k], chal[k] = v, v, v
    wit[3], rnd[3], chal[3] = 16, 0, 16
    wit[4, ::2], rnd[4, ::2], chal[4, ::2] = 16, 16, 16
    sat, rs, ch, _ = W.make_batch(5, 0, n, "U17")
    rs[:, :] = 16
    ch[:, :] = 16
    dense = np.full(44, 16, np.uint8)                 # every selector value 16; copy constraints of the test circuit
    dense[20:] = C[20:]
    for fast in (False, True, "wide"):
        impl = HostcheckImpl(oracle, fast=fast)
        for mode in ("generator9", "identity6") if fast else ("generator9", "identity6", "generator4"):
            g1s, g2 = util.SRS_MODES[mode](W)
            for circuit in (C, dense):
                ps.eq(f"extreme random {fast} {mode}", impl.plonk_prove_batch(circuit, g1s, g2, wit, rnd, chal),
                      oracle.plonk_prove_batch(circuit, g1s, g2, wit, rnd, chal, 8))
                ps.eq(f"extreme satisfying {fast} {mode}", impl.plonk_prove_batch(circuit, g1s, g2, sat, rs, ch),
                      oracle.plonk_prove_batch(circuit, g1s, g2, sat, rs, ch, 8))


def test_hostcheck_barrett_reductions_over_their_whole_range(hc):
    """red17 / red101 (multiply-high Barrett step, field.cuh) against `%` for EVERY x below the bounds the kernels rely on
    (2^28 resp. 2^26), and where above those bounds the formulas first fail."""
    import ctypes as C
    b17, b101 = C.c_uint32(), C.c_uint32()
    hc.lib.hc_check_barrett.restype = C.c_uint64
    assert hc.lib.hc_check_barrett(C.byref(b17), C.byref(b101)) == 0
    assert b17.value == 0 or b17.value >= 1 << 28
    assert b101.value == 0 or b101.value >= 1 << 26
    print("first failing x: red17", b17.value, "red101", b101.value)


def test_hostcheck_pairing_on_every_canonical_input(hc, oracle):
    """pairing17 (one doubling chain, shared squarings, f^600 = conj(f^5) f^95) against the oracle's pairing for EVERY
    canonically encoded G1 point (101 curve points + identity) and EVERY G2 byte pair below 101: 1 040 502 pairings."""
    import numpy as np
    pts = np.concatenate([ps.curve_points(), np.array([[0, 0, 1]], np.uint8)])
    q = np.stack(np.meshgrid(np.arange(101, dtype=np.uint8), np.arange(101, dtype=np.uint8), indexing="ij"), -1).reshape(-1, 2)
    P = np.repeat(pts, len(q), axis=0)
    Q = np.tile(q, (len(pts), 1))
    ps.eq("pairing, all canonical (P, Q)", hc.pairing(P, Q), oracle.pairing(P, Q, 8))


def test_hostcheck_wide_table_is_srs_eval_at_s_everywhere(hc, oracle, W):
    """EVERY entry of the one-look-up commitment table against the oracle's srs_eval_at_s: all 17^6 = 24 137 569
    six-coefficient polynomials over SRS rows 0..5, all 17^3 over rows 6..8 (generator SRS, n = 9)."""
    import numpy as np
    g1s, g2 = W.generator_srs(9)
    t6, t3 = hc.wide_tables(g1s)
    idx = np.arange(17 ** 6, dtype=np.int64)
    polys = np.empty((17 ** 6, 6), np.uint8)
    for k in range(6):
        polys[:, k] = idx % 17
        idx //= 17
    want, st = oracle.srs_eval_at_s(g1s, g2, polys, np.full(17 ** 6, 6, np.uint8), 8)
    assert (st == 0).all()
    ps.eq("T6 == srs_eval_at_s on every 6-coefficient polynomial", t6, want)
    idx = np.arange(17 ** 3, dtype=np.int64)
    hi = np.zeros((17 ** 3, 9), np.uint8)
    for k in range(3):
        hi[:, 6 + k] = idx % 17
        idx //= 17
    want3, st3 = oracle.srs_eval_at_s(g1s, g2, hi, np.full(17 ** 3, 9, np.uint8), 8)
    assert (st3 == 0).all()
    ps.eq("T3 == srs_eval_at_s on rows 6..8", t3, want3)


def test_hostcheck_wide_tables(oracle, W):
    """One-look-up commitments (T6, 17^6 entries): same eligibility as the pair tables, byte-identical output required."""
    import util
    hcw = HostcheckImpl(oracle, fast="wide")
    modes = [(m, f) for m, f in util.SRS_MODES.items()]
    ps.check_protocol(hcw, oracle, W, n=8000, modes=modes)
    ps.check_golden_transcript(hcw, W)
    ps.check_fiat_shamir(hcw, oracle, W, n=4000, modes=modes[3:4])
    ps.check_whole_curve_srs(hcw, oracle, W, n=2000, trials=3)
    ps.check_random_circuits(hcw, oracle, W, n=1500, circuits=4)


def test_hostcheck_fiat_shamir(oracle, W):
    """Fiat-Shamir mode of prove_one / the verifier kernels' challenge derivation (transcript.cuh) against oracle/fs_spec.inc"""
    import util
    ps.check_fiat_shamir(HostcheckImpl(oracle), oracle, W, n=6000)
    ps.check_fiat_shamir(HostcheckImpl(oracle, fast=True), oracle, W, n=6000, modes=[(m, f) for m, f in util.SRS_MODES.items()])
    ps.check_fiat_shamir_golden(HostcheckImpl(oracle), W)


def test_hostcheck_random_circuits(oracle, W):
    ps.check_random_circuits(HostcheckImpl(oracle), oracle, W, n=2000, circuits=6)
    ps.check_random_circuits(HostcheckImpl(oracle, fast=True), oracle, W, n=2000, circuits=6)


def test_hostcheck_whole_curve_srs(oracle, W):
    ps.check_whole_curve_srs(HostcheckImpl(oracle, fast=True), oracle, W, n=3000, trials=6)
    ps.check_whole_curve_srs(HostcheckImpl(oracle), oracle, W, n=2000, trials=2)


def test_hostcheck_log_verifier(oracle, W):
    """The table-path verifier (discrete logarithms mod 102 + two pairing tables): its premise exhaustively, then the same
    protocol / Fiat-Shamir / whole-curve-SRS / random-circuit suites as the Straus verifier, verdicts and pairing values."""
    import ctypes as C
    import util
    hcl = HostcheckImpl(oracle, fast="log")
    hcl.lib.hc_check_discrete_logs.restype = C.c_uint64
    assert hcl.lib.hc_check_discrete_logs() == 0
    modes = [(m, f) for m, f in util.SRS_MODES.items()]
    ps.check_protocol(hcl, oracle, W, n=8000, modes=modes)
    ps.check_golden_transcript(hcl, W)
    ps.check_fiat_shamir(hcl, oracle, W, n=4000, modes=modes)
    ps.check_whole_curve_srs(hcl, oracle, W, n=3000, trials=6)
    ps.check_random_circuits(hcl, oracle, W, n=1500, circuits=4)
