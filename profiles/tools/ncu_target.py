"""Launch, once each, the kernels whose ncu captures profiles/r2 holds.  Run plain first, then under ncu:

    python profiles/tools/ncu_target.py headline            # prove + verify of the bench step, 2^21 items (3 launches each)
    python profiles/tools/ncu_target.py families            # pairing, g1_mul<u8>, field_op<101>, field_pow<101>, generic and fast poly kernels, config2, gather
    ncu --set full --clock-control none --import-source on -k regex:'prove_kernel|verify_fast' -s 4 -c 2 -o gpurun_out/prof_headline python profiles/tools/ncu_target.py headline
Items per launch are printed so that profiles/tools/ncu_digest.py can be given --items.
"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from plonk_c_b200 import host, wire, workload as W

what = sys.argv[1] if len(sys.argv) > 1 else "headline"
dev = torch.device("cuda", 0)
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
lib = host.lib()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())

if what in ("headline", "headline_pair", "headline_packed"):
    import os
    if what == "headline_pair":
        os.environ["PB_WIDE_TABLES"] = "0"
        os.environ["PB_VERIFY_TABLES"] = "0"       # the configuration without context-sized tables: pair-table prover, Straus verifier
    n = 1 << 21
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    sets = []
    for b in range(3):      # three different batches, as the bench rotates its buffer sets
        batch = W.make_batch(2025 + b, 0, n, "U17")
        sets.append([T(x) for x in batch] + [T(wire.pack_inputs3(*batch))])
    proofs = torch.empty((n, 34), dtype=torch.uint8, device=dev)
    status = torch.empty(n, dtype=torch.uint8, device=dev)
    verdict = torch.empty(n, dtype=torch.uint8, device=dev)
    counts = torch.zeros(18, dtype=torch.int64, device=dev)
    for wit, rnd, chal, u, packed in sets:
        if what == "headline_packed":
            pk.prove_verify_packed_dev(packed, v3=True)
        else:
            # the bench step's call: prover launch, verifier launch (which also produces the batch's counters)
            host._check(lib.pb_plonk_prove_verify_tally_dev(pk._h, P(wit), P(rnd), P(chal), P(u), P(proofs), P(status), P(verdict), P(counts), C.c_size_t(n), sp, None))
    torch.cuda.synchronize()
    print(f"{what}: 3 x (prove, verify) launches of {n} items; completed {(status == 0).sum().item()}")
else:
    g1s, g2 = W.generator_srs(9)
    # config 4: 2^22 pairings
    n = 1 << 22
    ai, bi, sc = W.make_group_items(2025, 0, n)
    Pn = T(W.g1_subgroup_table()[ai])
    Q = host.g2_mul(T(np.tile(np.array([[36, 31]], np.uint8), (n, 1))), T(bi.astype(np.int64)))
    host.pairing(Pn, Q)
    print("pairing_kernel", n)
    # config 3: 2^24 scalar multiplications, one-byte scalars
    n = 1 << 24
    ai, bi, sc = W.make_group_items(2025, 0, n)
    host.g1_mul(T(g1s[ai % 10]), T(sc))
    print("g1_mul_kernel<unsigned char>", n)
    # family 1: gf_mul, gf_pow over 2^27 elements
    n = 1 << 27
    rng = np.random.default_rng(2025)
    A, B = T(rng.integers(0, 101, n, dtype=np.uint8)), T(rng.integers(1, 101, n, dtype=np.uint8))
    host.gf_mul(A, B)
    print("field_op_kernel<101>", n)
    host.gf_pow(A, B)
    print("field_pow_kernel<101>", n)
    del A, B
    # family 2, generic shapes (no fixed-shape fast path): 8x5 product, 13 / 4 division; then the config-2 shapes
    n = 1 << 22
    a = T(rng.integers(0, 17, (n, 8), dtype=np.uint8)); al = T(rng.integers(1, 9, n).astype(np.uint8))
    b = T(rng.integers(0, 17, (n, 5), dtype=np.uint8)); bl = T(rng.integers(1, 6, n).astype(np.uint8))
    host.poly_mul(a, al, b, bl)
    print("poly_binop_kernel (8 x 5)", n)
    num = T(rng.integers(0, 17, (n, 13), dtype=np.uint8)); nl = T(rng.integers(1, 14, n).astype(np.uint8))
    den = T(rng.integers(0, 17, (n, 4), dtype=np.uint8)); dl = T(rng.integers(1, 5, n).astype(np.uint8))
    host.poly_divide(num, nl, den, dl)
    print("poly_divide_kernel (13 / 4)", n)
    a6, b6, x, vals = [T(v) for v in W.make_poly_items(2025, 0, n)]
    six, five = T(np.full(n, 6, np.uint8)), T(np.full(n, 5, np.uint8))
    zh = T(np.tile(np.array([16, 0, 0, 0, 1], np.uint8), (n, 1)))
    prod, plen = host.poly_mul(a6, six, b6, six)
    print("poly_mul_fast_kernel<6, 6>", n)
    host.poly_divide(prod, plen, zh, five, sq=7, sr=4)
    print("poly_divide_fast_kernel<11, 5>", n)
    ctx0 = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    ctx0.poly_divide_zh(prod, plen)
    print("poly_divide_zh4_kernel<11>", n)
    host.poly_eval(a6, six, x)
    print("poly_eval_fast_kernel<6>", n)
    ctx = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    ctx.interpolate_at_h(vals)
    print("interpolate_kernel", n)
    ctx.config2_items(a6, b6, x, vals)
    print("config2_kernel", n)
    torch.cuda.synchronize()
