"""Regenerates tests/golden/*.npz from the UNMODIFIED reference compiled here (oracle/_ref).  Run in the
build container (needs /root/reference); the .npz files are committed and travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle._binding import load_ref  # noqa: E402
import plonk_c_b200.workload as W  # noqa: E402
import util  # noqa: E402

R = load_ref()
C = W.PLONK_TEST_CIRCUIT
N = 2048

# protocol: every SRS mode x both variants, seed 77
out = {}
for mode, mk in list(util.SRS_MODES.items()) + [("garbage10", lambda W: util.garbage_srs())]:
    g1s, g2 = mk(W)
    for var in ("U17", "NZ"):
        wit, rnd, chal, u = W.make_batch(77, 0, N, var)
        wit[0], rnd[0], chal[0], u[0] = W.GOLDEN_WITNESS[0], W.GOLDEN_RAND[0], W.GOLDEN_CHALLENGE[0], W.GOLDEN_U[0]
        proofs, status = R.plonk_prove_batch(C, g1s, g2, wit, rnd, chal)
        verdict, gt = R.plonk_verify_batch(C, g1s, g2, proofs, chal, u)
        k = f"{mode}_{var}"
        out[k + "_proofs"], out[k + "_status"], out[k + "_verdict"], out[k + "_gt"] = proofs, status, verdict, gt
    out[mode + "_g1s"], out[mode + "_g2"] = g1s, g2
    out[mode + "_vkey"] = R.verifier_key(C, g1s, g2)
np.savez_compressed(os.path.join(HERE, "protocol.npz"), **out)

# Fiat-Shamir mode (spec oracle/fs_spec.inc): the unmodified prover driven through the transcript, seed 78
out = {}
for mode, mk in list(util.SRS_MODES.items()) + [("garbage10", lambda W: util.garbage_srs())]:
    g1s, g2 = mk(W)
    seed = R.fs_seed(C, g1s, g2)                          # four 32-bit words as one int
    out[mode + "_seed"] = np.array([(seed >> (32 * k)) & 0xFFFFFFFF for k in range(4)], np.uint32)
    for var in ("U17", "NZ"):
        wit, rnd, _, _ = W.make_batch(78, 0, N, var)
        wit[0], rnd[0] = W.GOLDEN_WITNESS[0], W.GOLDEN_RAND[0]
        proofs, status, chal = R.plonk_prove_fs_batch(C, g1s, g2, wit, rnd)
        derived = R.fs_derive(R.fs_seed(C, g1s, g2), proofs)
        verdict, gt = R.plonk_verify_batch(C, g1s, g2, proofs, np.ascontiguousarray(derived[:, :5]), np.ascontiguousarray(derived[:, 5]))
        k = f"{mode}_{var}"
        out[k + "_proofs"], out[k + "_status"], out[k + "_chal"], out[k + "_verdict"], out[k + "_gt"] = proofs, status, chal, verdict, gt
np.savez_compressed(os.path.join(HERE, "fiat_shamir.npz"), **out)

# groups / pairing
a, b = util.g1_cases(N)
s = util.scalars_u64(N)
s2 = s.copy(); s2[s2 == 0] = 1
e = s.copy(); e[N // 2:] %= 20000
P, Q, sc = util.subgroup_points(W, R, N)
np.savez_compressed(
    os.path.join(HERE, "groups.npz"),
    g1_add=R.g1_op(0, a, b), g1_double=R.g1_op(1, a), g1_neg=R.g1_op(2, a), g1_mul=R.g1_mul(a, s),
    g1_on_curve=R.g1_is_on_curve(a), g2_add=R.g2_op(0, a[:, :2], b[:, :2]), g2_neg=R.g2_op(2, a[:, :2]),
    g2_mul=R.g2_mul(a[:, :2], s2), gtp_mul=R.gtp_mul(a[:, :2], b[:, :2]), gtp_pow=R.gtp_pow(a[:, :2], e),
    line=R.line(a, b), pairing_any=R.pairing(a, b[:, :2]), pairing_subgroup=R.pairing(P, Q),
    pairing_f5=R.pairing_f(5, a, b[:, :2]), g1_mul_subgroup=R.g1_mul(P, sc.astype(np.uint64)))

# polynomials
out = {}
for la, lb in ((6, 6), (11, 6), (16, 7), (11, 5), (22, 5), (10, 2), (3, 9)):
    pa, al, pb, bl = util.poly_cases(N, la, lb)
    k = f"{la}x{lb}"
    for op, nm in ((0, "add"), (1, "sub"), (2, "mul")):
        o, ol = R.poly_binop(op, pa, al, pb, bl, la + lb)
        out[f"{nm}_{k}"], out[f"{nm}_{k}_len"] = o, ol
    q, ql, r, rl, st = R.poly_divide(pa, al, pb, bl, la, lb)
    out[f"div_{k}_q"], out[f"div_{k}_ql"], out[f"div_{k}_r"], out[f"div_{k}_rl"], out[f"div_{k}_st"] = q, ql, r, rl, st
    x = pa[:, 0].copy()
    out[f"eval_{k}"] = R.poly_eval(pa, al, x)
vals = util.poly_cases(N, 4, 4)[0]
out["interp"], out["interp_len"] = R.interpolate_at_h(vals)
np.savez_compressed(os.path.join(HERE, "polys.npz"), **out)
print("golden fixtures written to", HERE)
