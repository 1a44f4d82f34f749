"""GPU parity tests: libplonk_b200.so through its C ABI against the CPU oracle, bit-exact.  Both boundary
flavours are exercised: the `_dev` entry points on torch CUDA tensors (device path) and the host-pointer
entry points on numpy arrays (host path).  Nothing here reads /root/reference."""
import numpy as np
import pytest

import parity_suite as ps
from impls import GpuImpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(host):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return GpuImpl(host, "device")


@pytest.fixture(scope="module")
def hostpath(host):
    return GpuImpl(host, "host")


def test_reference_known_answers(dev):
    ps.check_reference_known_answers(dev)


def test_reference_known_answers_host_path(hostpath):
    ps.check_reference_known_answers(hostpath)


def test_golden_transcript(dev, hostpath, W):
    ps.check_golden_transcript(dev, W)
    ps.check_golden_transcript(hostpath, W)


def test_fields(dev, oracle):
    ps.check_fields_exhaustive(dev, oracle)
    ps.check_fields_pow_all_bytes(dev, oracle)
    ps.check_fields_ragged(dev, oracle)


def test_fields_host_path(hostpath, oracle):
    ps.check_fields_ragged(hostpath, oracle)


def test_polys(dev, oracle):
    ps.check_polys(dev, oracle)


def test_polys_fast_shapes(dev, oracle):
    ps.check_polys_fast_shapes(dev, oracle)


def test_polys_golden(dev, hostpath):
    ps.check_polys_golden(dev)
    ps.check_polys_golden(hostpath)


def test_groups(dev, oracle):
    ps.check_groups(dev, oracle)


def test_pairing(dev, oracle, W):
    ps.check_pairing(dev, oracle, W)


def test_groups_golden(dev, hostpath, oracle, W):
    ps.check_groups_golden(dev, W, oracle)
    ps.check_groups_golden(hostpath, W, oracle)


def test_commitments(dev, oracle, W):
    ps.check_commitments(dev, oracle, W)


def test_protocol(dev, oracle, W):
    ps.check_protocol(dev, oracle, W)


def test_protocol_host_path(hostpath, oracle, W):
    ps.check_protocol(hostpath, oracle, W, n=20000, modes=[("generator9", lambda W: W.generator_srs(9)), ("identity6", lambda W: W.identity_srs(6))])


def test_protocol_golden(dev, hostpath, W):
    ps.check_protocol_golden(dev, W)
    ps.check_protocol_golden(hostpath, W)


def test_fiat_shamir(dev, oracle, W):
    """Fiat-Shamir mode: the transcript-driven prover and verifier kernels against
    oracle/fs_spec.inc, every SRS mode (fast and exact table paths)."""
    ps.check_fiat_shamir(dev, oracle, W, n=20000)


def test_fiat_shamir_host_path(hostpath, oracle, W):
    ps.check_fiat_shamir(hostpath, oracle, W, n=70000, modes=[("generator9", lambda W: W.generator_srs(9)), ("identity6", lambda W: W.identity_srs(6))])


def test_fiat_shamir_golden(dev, hostpath, W):
    ps.check_fiat_shamir_golden(dev, W)
    ps.check_fiat_shamir_golden(hostpath, W)


def test_fiat_shamir_exact_path_forced(host, oracle, W):
    import os
    os.environ["PB_FORCE_EXACT"] = "1"
    try:
        ps.check_fiat_shamir(GpuImpl(host, "device"), oracle, W, n=20000, modes=[("generator9", lambda W: W.generator_srs(9))])
    finally:
        del os.environ["PB_FORCE_EXACT"]


def test_setup_constants(host, oracle, W):
    """plonk_new + the witness-independent part of plonk_prove, as computed by the context's setup kernels."""
    for mk in (W.identity_srs, W.generator_srs):
        g1s, g2 = mk(9)
        pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
        ps.eq("plonk_new", pk.setup_dump(), oracle.plonk_setup_dump())
        d = pk.circuit_dump()
        assert d["sigma"].tolist() == [[2, 8, 15, 3], [1, 4, 16, 12], [13, 9, 5, 14]]          # plonk-test.c:105-112
        assert d["s_sigma"].tolist() == [[7, 13, 10, 6], [4, 0, 13, 1], [6, 7, 3, 14]]
        assert d["l1"].tolist() == [13, 13, 13, 13]
        ps.eq("verifier key", pk.verifier_key(), oracle.verifier_key(W.PLONK_TEST_CIRCUIT, g1s, g2))
        tab = pk.srs_table()
        want = np.stack([oracle.g1_mul(np.tile(g1s[i:i + 1], (17, 1)), np.arange(17, dtype=np.uint64)) for i in range(10)])
        ps.eq("fixed-base table = g1_mul(g1s[i], c)", tab, want)
        sat = pk.constraints_satisfy(W.satisfying_witnesses())
        assert sat.all()
        bad = W.satisfying_witnesses().copy()
        bad[:, 11] = (bad[:, 11] + 1) % 17
        assert not pk.constraints_satisfy(bad).any()


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 1000, 4097])
def test_ragged_batch_sizes(host, oracle, W, n):
    """Empty, single and non-multiple-of-block batches through prove, verify and the fused prove_verify."""
    import torch
    g1s, g2 = W.generator_srs(9)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
    wit, rnd, chal, u = W.make_batch(31, 1000, n, "U17")
    d = [torch.from_numpy(x).cuda() for x in (wit, rnd, chal, u)]
    proofs, status, verdict = pk.prove_verify(*d)
    hp, hs, hv = pk.prove_verify(wit, rnd, chal, u)                     # host-pointer pipeline
    if n == 0:
        assert proofs.shape == (0, 34) and hp.shape == (0, 34)
        return
    rp, rs = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    rv, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, rp, chal, u)
    rv = np.where(rs == 0, rv, 0xFF).astype(np.uint8)
    ps.eq("device path", (proofs.cpu().numpy(), status.cpu().numpy(), verdict.cpu().numpy()), (rp, rs, rv))
    ps.eq("host path", (hp, hs, hv), (rp, rs, rv))


def test_bad_input_bytes_are_flagged_not_undefined(host, W):
    """Bytes outside [0,17) are not a reference path (HF is 'always kept in range', hf.h:11-14): status 254."""
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    wit, rnd, chal, u = W.make_batch(3, 0, 512)
    wit[5, 3] = 17
    rnd[77, 0] = 255
    chal[300, 4] = 200
    proofs, status = pk.prove(wit, rnd, chal)
    assert status[5] == 254 and status[77] == 254 and status[300] == 254
    assert (status[[5, 77, 300]] == 254).all() and not proofs[[5, 77, 300]].any()
    assert (np.delete(status, [5, 77, 300]) != 254).all()


def test_misaligned_device_pointer_is_rejected(host, W):
    import torch
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    wit, rnd, chal, u = W.make_batch(3, 0, 64)
    big = torch.zeros(64 * 12 + 1, dtype=torch.uint8, device="cuda")
    view = big[1:].view(64, 12)
    with pytest.raises(host.PlonkB200Error) as e:
        pk.prove(view, torch.from_numpy(rnd).cuda(), torch.from_numpy(chal).cuda())
    assert e.value.code == host.PB_ERR_ARG


def test_protocol_exact_path_forced(host, oracle, W):
    """PB_FORCE_EXACT=1 makes a context take the sequential (any-SRS) prover and verifier even for a canonical SRS:
    both code paths must agree with the oracle on the same inputs."""
    import os
    os.environ["PB_FORCE_EXACT"] = "1"
    try:
        impl = GpuImpl(host, "device")
        ps.check_protocol(impl, oracle, W, n=30000, modes=[("generator9", lambda W: W.generator_srs(9)), ("identity6", lambda W: W.identity_srs(6))])
        ps.check_golden_transcript(impl, W)
    finally:
        del os.environ["PB_FORCE_EXACT"]


def test_protocol_pair_tables_path(host, oracle, W):
    """PB_WIDE_TABLES=0 keeps a context on the pair-table prover (5.8 KB of tables in shared memory instead of the 48 MB
    one-look-up table): the smaller-footprint fast path must stay byte-identical too."""
    import os
    os.environ["PB_WIDE_TABLES"] = "0"
    try:
        impl = GpuImpl(host, "device")
        modes = [("generator9", lambda W: W.generator_srs(9)), ("generator6", lambda W: W.generator_srs(6)), ("identity6", lambda W: W.identity_srs(6))]
        ps.check_protocol(impl, oracle, W, n=30000, modes=modes)
        ps.check_fiat_shamir(impl, oracle, W, n=10000, modes=modes[:1])
        ps.check_golden_transcript(impl, W)
    finally:
        del os.environ["PB_WIDE_TABLES"]


def test_protocol_straus_verifier_path(host, oracle, W):
    """PB_VERIFY_TABLES=0 keeps a context on the Straus + Miller-loop verifier kernel (no discrete-logarithm / pairing
    tables): the arithmetic path must stay identical to the oracle too, verdicts and pairing values."""
    import os
    os.environ["PB_VERIFY_TABLES"] = "0"
    try:
        impl = GpuImpl(host, "device")
        modes = [("generator9", lambda W: W.generator_srs(9)), ("generator6", lambda W: W.generator_srs(6)), ("identity6", lambda W: W.identity_srs(6))]
        ps.check_protocol(impl, oracle, W, n=30000, modes=modes)
        ps.check_fiat_shamir(impl, oracle, W, n=10000, modes=modes[:1])
        ps.check_whole_curve_srs(impl, oracle, W, n=3000, trials=3)
        ps.check_golden_transcript(impl, W)
    finally:
        del os.environ["PB_VERIFY_TABLES"]


def test_host_pipeline_pinned_matches_pageable(host, W):
    """Host-pointer prove+verify with pinned buffers against the same call with pageable buffers, explicit and
    Fiat-Shamir mode, a ragged size spanning many pipeline chunks."""
    import torch
    n = 300007
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    wit, rnd, chal, u = W.make_batch(321, 0, n, "U17")
    want = pk.prove_verify(wit, rnd, chal, u)                      # numpy (pageable) in and out
    pin_in = [torch.from_numpy(x).pin_memory().numpy() for x in (wit, rnd, chal, u)]
    pin_out = [torch.full((n, 34), 7, dtype=torch.uint8).pin_memory().numpy(), torch.full((n,), 7, dtype=torch.uint8).pin_memory().numpy(),
               torch.full((n,), 7, dtype=torch.uint8).pin_memory().numpy()]
    pk.prove_verify_into(*pin_in, *pin_out)
    ps.eq("pinned outputs, explicit challenges", tuple(pin_out), want)
    want_fs = pk.prove_verify_fs(wit, rnd)
    for o in pin_out:
        o[...] = 9
    pk.prove_verify_fs_into(pin_in[0], pin_in[1], *pin_out)
    ps.eq("pinned outputs, Fiat-Shamir", tuple(pin_out), want_fs)
    proofs, status = pk.prove(wit, rnd, chal)
    ps.eq("prove only", (proofs, status), want[:2])
    ps.eq("prove only, Fiat-Shamir, pipelined (no challenge read-back)", pk.prove_fs(wit, rnd), want_fs[:2])
    ps.eq("prove only, Fiat-Shamir, with challenge read-back", pk.prove_fs(wit, rnd, want_challenges=True)[:2], want_fs[:2])


def test_prove_verify_fused_matches_separate_calls(host, W):
    """pb_plonk_prove_verify (prover-fed dense list) == pb_plonk_prove + pb_plonk_verify on the completed proofs."""
    import torch
    for mk in (W.generator_srs, W.identity_srs):
        pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *mk(9))
        wit, rnd, chal, u = W.make_batch(123, 0, 50000, "U17")
        d = [torch.from_numpy(x).cuda() for x in (wit, rnd, chal, u)]
        proofs, status, verdict = pk.prove_verify(*d)
        p2, s2 = pk.prove(d[0], d[1], d[2])
        v2 = pk.verify(p2, d[2], d[3])
        v2 = torch.where(s2 == 0, v2, torch.full_like(v2, 0xFF))
        assert torch.equal(proofs, p2) and torch.equal(status, s2) and torch.equal(verdict, v2)


def test_invalid_copy_type_circuit(host, oracle, W):
    """A COPY_OF.type outside {A,B,C}: the reference exits with "Invalid copy_of type" (plonk.h:155-157) for every
    satisfied witness -> status 3 -- after the constraints assert (status 1 wins for unsatisfied witnesses)."""
    circuit = W.PLONK_TEST_CIRCUIT.copy()
    circuit[20 + 2] = 7                       # c_a[2].type
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(17, 0, 4000, "U17")
    wit[::7, 8] = (wit[::7, 8] + 1) % 17      # every 7th witness violates gate 0
    want = oracle.plonk_prove_batch(circuit, g1s, g2, wit, rnd, chal)
    pk = host.Plonk(circuit, g1s, g2)
    got = pk.prove(wit, rnd, chal)
    ps.eq("invalid copy type", got, want)
    assert set(np.unique(got[1])) == {1, 3}


def test_abi_argument_errors(host, W):
    g1s, g2 = W.generator_srs(9)
    bad = W.PLONK_TEST_CIRCUIT.copy()
    bad[24] = 0                                # a 1-based copy index of 0 underflows in the reference (plonk.h:144): rejected
    with pytest.raises(host.PlonkB200Error) as e:
        host.Plonk(bad, g1s, g2)
    assert e.value.code == host.PB_ERR_ARG
    bad = W.PLONK_TEST_CIRCUIT.copy()
    bad[3] = 17                                # selector byte outside F17
    with pytest.raises(host.PlonkB200Error):
        host.Plonk(bad, g1s, g2)
    g = g1s.copy()
    g[0, 0] = 101                              # SRS coordinate outside F101
    with pytest.raises(host.PlonkB200Error):
        host.Plonk(W.PLONK_TEST_CIRCUIT, g, g2)
    a = np.zeros((4, 70), np.uint8)            # longer than PB_POLY_MAX
    with pytest.raises(host.PlonkB200Error):
        host.poly_mul(a, np.full(4, 70, np.uint8), a, np.full(4, 70, np.uint8))
    with pytest.raises(host.PlonkB200Error):
        host.field_op(19, 0, np.zeros(16, np.uint8), np.zeros(16, np.uint8))


def test_config2_fused(host, oracle, W):
    """pb_config2_items_dev == the four separate reference functions composed by the oracle, incl. zero / short products."""
    import torch
    n = 50021
    a, b, x, vals = W.make_poly_items(5, 0, n)
    a[:500] = 0
    b[500:900, 1:] = 0
    a[900:1300, 3:] = 0
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    got = pk.config2_items(*[torch.from_numpy(v).cuda() for v in (a, b, x, vals)])
    got = [g.cpu().numpy() for g in got]
    six, five = np.full(n, 6, np.uint8), np.full(n, 5, np.uint8)
    zh = np.tile(np.array([16, 0, 0, 0, 1], np.uint8), (n, 1))
    prod, plen = oracle.poly_binop(2, a, six, b, six, 11)
    quot, qlen, rem, rlen, st = oracle.poly_divide(prod, plen, zh, five, 7, 4)
    ev = oracle.poly_eval(a, six, x)
    ip, il = oracle.interpolate_at_h(vals)
    assert not st.any()
    ps.eq("config2 fused", tuple(got), (prod, plen, quot, qlen, rem, rlen, ev, ip, il))


def test_random_circuits(dev, oracle, W):
    ps.check_random_circuits(dev, oracle, W)


def test_field_new(host, oracle):
    """hf_new / gf_new on negatives and on +-1.7e18 (hf-test.c:231-248, gf-test.c:44-88)."""
    import torch
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.integers(-2**62, 2**62, 5000), np.arange(-300, 300),
                        np.array([-(17 ** 14), 17 ** 14, -1, 0, 1, 2**63 - 1, -2**63, 1700000000000000000, -1700000000000000000])]).astype(np.int64)
    for field, ofn in ((17, oracle.hf_new), (101, oracle.gf_new)):
        want = np.array([ofn(int(x)) for x in v], np.uint8)
        ps.eq(f"new{field} host path", host.field_new(field, v), want)
        ps.eq(f"new{field} device path", host.field_new(field, torch.from_numpy(v).cuda()).cpu().numpy(), want)


def test_whole_curve_srs(dev, oracle, W):
    ps.check_whole_curve_srs(dev, oracle, W)


# ---------------------------------------------------------------- compact output, packed wire v2, seeded mode
def _oracle_pv(oracle, W, circuit, g1s, g2, wit, rnd, chal, u):
    rp, rs = oracle.plonk_prove_batch(circuit, g1s, g2, wit, rnd, chal)
    rv, _ = oracle.plonk_verify_batch(circuit, g1s, g2, rp, chal, u)
    return rp, rs, np.where(rs == 0, rv, 0xFF).astype(np.uint8)


@pytest.mark.parametrize("n", [0, 1, 127, 129, 4097, 300007])
def test_compact_and_packed_outputs(host, oracle, W, n):
    """pb_plonk_prove_verify_compact / _packed (host-pointer pipelines) against the oracle: the dense list is exactly the
    completed proofs in item order, the packed records are the wire.py packing of them, sv the packed status/verdict."""
    from plonk_c_b200 import wire
    for mode in ("generator9", "identity6", "garbage"):
        import util
        g1s, g2 = util.garbage_srs() if mode == "garbage" else util.SRS_MODES[mode](W)
        if mode != "generator9" and n > 5000:
            continue
        pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
        wit, rnd, chal, u = W.make_batch(55, 7, n, "U17")
        if n > 100:
            wit[5, 3] = 17          # a byte outside F17: status 254 in both formats
        dense, status, verdict = pk.prove_verify_compact(wit, rnd, chal, u)
        packed_in = wire.pack_inputs(wit, rnd, chal, u)
        pp, sv = pk.prove_verify_packed(packed_in)
        if n == 0:
            assert dense.shape == (0, 34) and pp.shape == (0, 22)
            continue
        if n > 100:
            wit_ok = wit.copy(); wit_ok[5] = 0
            rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit_ok, rnd, chal, u)
            rp[5] = 0; rs[5] = 254; rv[5] = 0xFF
        else:
            rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
        ps.eq(f"compact {mode}", (dense, status, verdict), (rp[rs == 0], rs, rv))
        ps.eq(f"compact {mode}: scatter", host.wire_scatter_proofs(dense, status), rp)
        ps.eq(f"packed {mode}", (pp, sv), (wire.pack_proofs(rp[rs == 0]), wire.make_sv(rs, rv)))
        ps.eq(f"packed {mode}: unpack", wire.scatter_proofs(wire.unpack_proofs(pp), wire.split_sv(sv)[0]), rp)


def test_packed_device_path_and_gather(host, oracle, W):
    """pb_plonk_prove_verify_packed_dev and pb_gather_completed_dev on device buffers; non-encodings are flagged."""
    import torch
    from plonk_c_b200 import wire
    n = 70001
    g1s, g2 = W.generator_srs(9)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
    wit, rnd, chal, u = W.make_batch(56, 0, n, "U17")
    packed = wire.pack_inputs(wit, rnd, chal, u)
    packed[11] = 0xFF                                   # word >= 17^7
    packed[12].view("<u4")[3] += 16 * 17 ** 6          # spare digit != 0
    pp, cnt, sv = pk.prove_verify_packed_dev(torch.from_numpy(packed).cuda())
    rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
    for i in (11, 12):
        rp[i] = 0; rs[i] = 254; rv[i] = 0xFF
    k = int(cnt.item())
    assert k == int((rs == 0).sum())
    ps.eq("packed dev", (pp[:k].cpu().numpy(), sv.cpu().numpy()), (wire.pack_proofs(rp[rs == 0]), wire.make_sv(rs, rv)))
    d = [torch.from_numpy(x).cuda() for x in (wit, rnd, chal)]
    proofs, status = pk.prove(*d)
    dense, cnt2 = host.gather_completed(proofs, status)
    k2 = int(cnt2.item())
    p_h, s_h = proofs.cpu().numpy(), status.cpu().numpy()
    ps.eq("gather", dense[:k2].cpu().numpy(), p_h[s_h == 0])
    # all-failed and all-completed extremes of the dense list
    st0 = torch.zeros_like(status)
    dense, cnt3 = host.gather_completed(proofs, st0)
    assert int(cnt3.item()) == n and torch.equal(dense, proofs)
    st1 = torch.full_like(status, 8)
    _, cnt4 = host.gather_completed(proofs, st1)
    assert int(cnt4.item()) == 0


def test_packed_exact_path_forced(host, oracle, W):
    import os
    from plonk_c_b200 import wire
    os.environ["PB_FORCE_EXACT"] = "1"
    try:
        n = 20000
        g1s, g2 = W.generator_srs(9)
        pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
        wit, rnd, chal, u = W.make_batch(57, 0, n, "NZ")
        pp, sv = pk.prove_verify_packed(wire.pack_inputs(wit, rnd, chal, u))
        rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
        ps.eq("packed, exact kernels", (pp, sv), (wire.pack_proofs(rp[rs == 0]), wire.make_sv(rs, rv)))
    finally:
        del os.environ["PB_FORCE_EXACT"]


def test_seeded_mode(host, oracle, W):
    """Seeded mode: the device-generated stream equals workload.make_batch, and the counters equal the tally of the
    oracle's results on those items (both variants, a start offset, a count that is not a multiple of anything)."""
    from plonk_c_b200 import shard, wire
    g1s, g2 = W.generator_srs(9)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
    for seed, start, n, variant in ((2025, 0, 50000, "U17"), (99, 2**33 + 11, 30011, "NZ")):
        got = [t.cpu().numpy() for t in pk.synth_batch(seed, start, n, variant)]
        wit, rnd, chal, u = W.make_batch(seed, start, n, variant)
        ps.eq(f"device generator {variant}", tuple(got), (wit, rnd, chal, u, wire.pack_inputs(wit, rnd, chal, u)))
        counts = pk.prove_verify_seeded(seed, start, n, variant)
        rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
        ps.eq(f"seeded counters {variant}", counts, shard.tally_host(rp, rs, rv))
    # a count spanning several internal chunks: compared with the struct path on the GPU itself
    import torch
    n = (1 << 21) + 12345
    counts = pk.prove_verify_seeded(5, 1000, n, "U17")
    wit, rnd, chal, u = W.make_batch(5, 1000, n, "U17")
    proofs, status, verdict = pk.prove_verify(*[torch.from_numpy(x).cuda() for x in (wit, rnd, chal, u)])
    ps.eq("seeded counters, 2^21 + 12345 items", counts, shard.tally_host(proofs.cpu().numpy(), status.cpu().numpy(), verdict.cpu().numpy()))


def test_output_buffers_are_not_coerced(host, W):
    """a non-contiguous or wrongly typed output buffer must raise, not be silently copied."""
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    n = 256
    wit, rnd, chal, u = W.make_batch(3, 0, n)
    good = [np.empty((n, 34), np.uint8), np.empty(n, np.uint8), np.empty(n, np.uint8)]
    pk.prove_verify_into(wit, rnd, chal, u, *good)
    with pytest.raises(ValueError):
        pk.prove_verify_into(wit, rnd, chal, u, np.empty((n, 68), np.uint8)[:, ::2], good[1], good[2])
    with pytest.raises(ValueError):
        pk.prove_verify_into(wit, rnd, chal, u, good[0], np.empty(n, np.int32), good[2])
    with pytest.raises(ValueError):
        pk.prove_verify_into(wit, rnd, chal, u, good[0], good[1], np.empty(n + 1, np.uint8))


def test_compact_and_packed_many_chunks(host, W):
    """A batch long enough to wrap the pipeline's ring of buffer sets (11 chunks over 6 slots): compact and packed outputs
    against the struct-array pipeline (itself checked against the oracle above)."""
    from plonk_c_b200 import wire
    n = (1 << 21) + 77
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    wit, rnd, chal, u = W.make_batch(58, 0, n, "U17")
    proofs, status, verdict = pk.prove_verify(wit, rnd, chal, u)
    dense, s2, v2 = pk.prove_verify_compact(wit, rnd, chal, u)
    ps.eq("compact, many chunks", (dense, s2, v2), (proofs[status == 0], status, verdict))
    pp, sv = pk.prove_verify_packed(wire.pack_inputs(wit, rnd, chal, u))
    ps.eq("packed, many chunks", (pp, sv), (wire.pack_proofs(proofs[status == 0]), wire.make_sv(status, verdict)))
    # and twice in a row on the same context (slot reuse across calls), with a smaller batch in between
    pk.prove_verify_compact(wit[:1000], rnd[:1000], chal[:1000], u[:1000])
    dense3, s3, v3 = pk.prove_verify_compact(wit, rnd, chal, u)
    ps.eq("compact, second call", (dense3, s3, v3), (dense, s2, v2))


def test_prove_verify_tally_fused(host, oracle, W):
    """pb_plonk_prove_verify_tally_dev: outputs as pb_plonk_prove_verify_dev, counters as pb_tally_dev of those outputs --
    with the table-path verifier (counters from its epilogue), with the arithmetic verifier and on the exact path (separate
    pass), for a ragged size, twice into the same counters."""
    import os
    import torch
    from plonk_c_b200 import shard
    n = 70001
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(71, 5, n, "U17")
    wit[17, 2] = 200                                     # a bad input byte: status 254 lands in the histogram's last bin
    u[33] = 99                                           # a bad verifier scalar: verdict 3
    d = [torch.from_numpy(x).cuda() for x in (wit, rnd, chal, u)]
    for env in ({}, {"PB_VERIFY_TABLES": "0"}, {"PB_FORCE_EXACT": "1"}):
        os.environ.update(env)
        try:
            pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
            counts = torch.zeros(shard.N_COUNTERS, dtype=torch.int64, device="cuda")
            p, s, v = pk.prove_verify_tally_dev(*d, counts)
            p2, s2, v2 = pk.prove_verify(*d)
            ps.eq(f"fused tally outputs {env}", tuple(t.cpu().numpy() for t in (p, s, v)), tuple(t.cpu().numpy() for t in (p2, s2, v2)))
            want = shard.tally_host(p2.cpu().numpy(), s2.cpu().numpy(), v2.cpu().numpy())
            ps.eq(f"fused tally counters {env}", counts.cpu().numpy(), want)
            pk.prove_verify_tally_dev(*d, counts)
            ps.eq(f"fused tally accumulates {env}", counts.cpu().numpy(), 2 * want)
            assert int(s.cpu()[17]) == 254 and int(v.cpu()[17]) == 0xFF and (int(s.cpu()[33]) != 0 or int(v.cpu()[33]) == 3)
        finally:
            for key in env:
                del os.environ[key]


def test_packed_v3(host, oracle, W):
    """Packed wire v3 (14 B in, 12 B per completed proof: points as curve indices) against the oracle, host and device
    paths, in every table mode; bad records flagged; an SRS off the curve is refused."""
    import os
    import torch
    import util
    from plonk_c_b200 import wire
    n = 70001
    for mode, env in (("generator9", {}), ("identity6", {}), ("generator9", {"PB_WIDE_TABLES": "0"}), ("generator9", {"PB_FORCE_EXACT": "1"})):
        os.environ.update(env)
        try:
            g1s, g2 = util.SRS_MODES[mode](W)
            pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
            wit, rnd, chal, u = W.make_batch(61, 3, n, "U17")
            packed = wire.pack_inputs3(wit, rnd, chal, u)
            packed[11] = 0xFF                                                   # 29-bit fields >= 17^7
            packed[12, 12:] = 0xFF; packed[12, 3] |= 0xE0; packed[12, 7] |= 0xE0; packed[12, 11] |= 0xE0     # G >= 17^6
            rp, rs, rv = _oracle_pv(oracle, W, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
            for i in (11, 12):
                rp[i] = 0; rs[i] = 254; rv[i] = 0xFF
            want = (wire.pack_proofs3(rp[rs == 0]), wire.make_sv(rs, rv))
            pp, sv = pk.prove_verify_packed3(packed)
            ps.eq(f"packed v3 host {mode} {env}", (pp, sv), want)
            ps.eq(f"packed v3 {mode}: unpack", wire.scatter_proofs(wire.unpack_proofs3(pp), wire.split_sv(sv)[0]), rp)
            ppd, cnt, svd = pk.prove_verify_packed_dev(torch.from_numpy(packed).cuda(), v3=True)
            k = int(cnt.item())
            ps.eq(f"packed v3 dev {mode} {env}", (ppd[:k].cpu().numpy(), svd.cpu().numpy()), want)
        finally:
            for key in env:
                del os.environ[key]
    # many chunks (wraps the ring of buffer sets) against the v2 path on the GPU itself, and n = 0
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    n = (1 << 21) + 77
    batch = W.make_batch(62, 0, n, "U17")
    pp2, sv2 = pk.prove_verify_packed(wire.pack_inputs(*batch))
    pp3, sv3 = pk.prove_verify_packed3(wire.pack_inputs3(*batch))
    ps.eq("packed v3, many chunks", (wire.unpack_proofs3(pp3), sv3), (wire.unpack_proofs(pp2), sv2))
    e = pk.prove_verify_packed3(np.empty((0, 14), np.uint8))
    assert e[0].shape == (0, 12) and e[1].shape == (0,)
    # an SRS with points off the curve has no v3 encoding for its commitments
    bad = host.Plonk(W.PLONK_TEST_CIRCUIT, *util.garbage_srs())
    with pytest.raises(host.PlonkB200Error):
        bad.prove_verify_packed3(wire.pack_inputs3(*W.make_batch(1, 0, 256, "U17")))


def test_pipeline_ring_reuse_with_small_chunks():
    """The ring of buffer sets wraps only for batches of more than eight chunks (4 M proofs at the default chunk size).  The
    chunk size is read once per process, so a child process runs the many-chunk comparisons with 2^15-item chunks: 65 chunks
    over 8 slots, every host-pointer format against the struct path."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PB_PIPE_CHUNK="32768")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-x", "-q", "-m", "gpu", "-k",
                        "many_chunks or pinned_matches_pageable or packed_v3", "-p", "no:cacheprovider"],
                       env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_srs_eval_raw_and_satisfy_rows(host, oracle, W):
    """The context-free entry points behind the drop-in srs_eval_at_s / constraints_satisfy: one launch each, exactly the
    reference's loops -- including an untrimmed polynomial over a garbage SRS, where a trailing zero term changes the result."""
    import torch
    import util
    rng = np.random.default_rng(21)
    n = 20000
    for g1s, _ in (W.generator_srs(9), util.garbage_srs(10)):
        polys = rng.integers(0, 17, (n, 8), dtype=np.uint8)
        polys[rng.random((n, 8)) < 0.3] = 0
        plen = rng.integers(1, 9, n).astype(np.uint8)
        pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, np.array([36, 31, 90, 82], np.uint8))
        want = pk.srs_eval_at_s(polys, plen)                               # the table-based kernel, checked against the oracle elsewhere
        ps.eq("raw == context", host.srs_eval_at_s_raw(g1s, polys, plen), want)
        ps.eq("raw == context (device)", tuple(t.cpu().numpy() for t in host.srs_eval_at_s_raw(g1s, torch.from_numpy(polys).cuda(), torch.from_numpy(plen).cuda())), want)
        # untrimmed: the reference's loop over POLY.len terms = the oracle's term-by-term sum
        pts = np.zeros((n, 3), np.uint8); pts[:, 2] = 1
        for k in range(8):
            term = oracle.g1_mul(np.tile(g1s[k:k + 1], (n, 1)), polys[:, k].astype(np.uint64))
            nxt = oracle.g1_op(0, pts, term)
            live = plen > k
            pts[live] = nxt[live]
        got, st = host.srs_eval_at_s_raw(g1s, polys, plen, trim=False)
        assert not st.any()
        ps.eq("untrimmed loop", got, pts)
    q = rng.integers(0, 17, (5, 7), dtype=np.uint8)
    q[2] = 16                                                            # q_o = -1: a row holds iff c = q_l a + q_r b + q_m a b + q_c
    a, b, c = (rng.integers(0, 17, (n, 7), dtype=np.uint8) for _ in range(3))
    rhs = (q[0] * a.astype(np.int64) + q[1] * b + q[3] * (a.astype(np.int64) * b) + q[4]) % 17
    c[::3] = rhs[::3]                                                    # every third item satisfies all seven rows
    c[1::3, :4] = rhs[1::3, :4]                                          # another third fails first at row 4 or later
    lhs = (rhs + 16 * c.astype(np.int64)) % 17
    want = np.where((lhs != 0).any(axis=1), (lhs != 0).argmax(axis=1), -1).astype(np.int32)
    ps.eq("constraints_satisfy_rows", host.constraints_satisfy_rows(q, a, b, c), want)
    assert (want == -1).sum() > 1000 and (want >= 0).sum() > 1000


def test_config2_host_path(host, W):
    """pb_config2_items with host pointers (the generic chunked pipeline, several chunks) == the device path."""
    import torch
    n = (1 << 20) + 12345
    a, b, x, vals = W.make_poly_items(6, 0, n)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    dev = [g.cpu().numpy() for g in pk.config2_items(*[torch.from_numpy(v).cuda() for v in (a, b, x, vals)])]
    ps.eq("config2 host path", tuple(pk.config2_items(a, b, x, vals)), tuple(dev))
    # and the other families through the same pipeline at a size that spans chunks: g1_mul, pairing
    ai, bi, sc = W.make_group_items(3, 0, 3 * (1 << 20) + 77)
    P = W.g1_subgroup_table()[ai]
    want = host.g1_mul(torch.from_numpy(P).cuda(), torch.from_numpy(sc).cuda()).cpu().numpy()
    ps.eq("g1_mul host path", host.g1_mul(P, sc), want)


@pytest.mark.parametrize("shape", [(8, 5), (6, 6), (11, 6)])      # a generic shape and two register-resident fast shapes
def test_adversarial_bytes_in_poly_entry_points(host, oracle, W, shape):
    """per-item bytes the ABI cannot trust.  A length above the row stride, a coefficient byte above 16, a shift that
    does not fit the output row are REPORTED (olen 0 / status 2 / 0xFF) -- never computed on, never read or written out of
    bounds -- and the valid items of the same batch are untouched.  Device and host paths, generic and fast kernels."""
    import torch
    import util
    sa, sb = shape
    n = 4099
    a, al, b, bl = util.poly_cases(n, sa, sb, seed=91)
    bad_len = np.arange(0, n, 7)
    bad_coef = np.arange(3, n, 11)
    bad_zero = np.arange(5, n, 13)
    a2, al2, b2, bl2 = a.copy(), al.copy(), b.copy(), bl.copy()
    al2[bad_len] = 200                                              # far above the stride (and above PB_POLY_MAX)
    b2[bad_coef, 0] = 17
    bl2[bad_coef] = np.maximum(bl2[bad_coef], 1)
    al2[bad_zero] = 0
    bad = np.zeros(n, bool); bad[bad_len] = True; bad[bad_coef] = True; bad[bad_zero] = True
    for op in (host.POLY_MUL, host.POLY_ADD, host.POLY_SUB):
        so = sa + sb - 1 if op == host.POLY_MUL else max(sa, sb)
        want, wl = oracle.poly_binop(op, a, al, b, bl, so)
        for path in ("host", "device"):
            args = (a2, al2, b2, bl2) if path == "host" else tuple(torch.from_numpy(v).cuda() for v in (a2, al2, b2, bl2))
            got, gl = host.poly_binop(op, *args, so)
            got, gl = (np.asarray(got.cpu()) if path == "device" else got), (np.asarray(gl.cpu()) if path == "device" else gl)
            assert (gl[bad] == 0).all() and not got[bad].any(), (op, path)
            ps.eq(f"binop {op} {path}: valid items", (got[~bad], gl[~bad]), (want[~bad], wl[~bad]))
    # poly_eval: 0xFF for an invalid row or an x byte above 16
    x = np.random.default_rng(5).integers(0, 17, n, dtype=np.uint8)
    x2 = x.copy(); x2[bad_zero] = 99
    a3 = a.copy(); a3[bad_coef, 0] = 255
    al3 = np.maximum(al, 1); al3[bad_len] = sa + 1
    got = host.poly_eval(torch.from_numpy(a3).cuda(), torch.from_numpy(al3).cuda(), torch.from_numpy(x2).cuda()).cpu().numpy()
    assert (got[bad] == 0xFF).all()
    ps.eq("eval: valid items", got[~bad], oracle.poly_eval(a, np.maximum(al, 1), x)[~bad])
    # poly_divide: status 2
    al_div = al.copy(); al_div[bad_len] = 200                       # (a zero numerator length is a valid input of poly_divide)
    quot, ql, rem, rl, st = host.poly_divide(torch.from_numpy(a2).cuda(), torch.from_numpy(al_div).cuda(),
                                             torch.from_numpy(b2).cuda(), torch.from_numpy(bl2).cuda())
    st = st.cpu().numpy()
    assert (st[bad_len] == 2).all() and (st[bad_coef] == 2).all() and not quot.cpu().numpy()[st == 2].any() and not rem.cpu().numpy()[st == 2].any()
    wq, wql, wr, wrl, wst = oracle.poly_divide(a, al, b, bl, sa, max(sb - 1, 1))
    ok = ~(np.isin(np.arange(n), bad_len) | np.isin(np.arange(n), bad_coef))
    ps.eq("divide: valid items", (quot.cpu().numpy()[ok], ql.cpu().numpy()[ok], rem.cpu().numpy()[ok], rl.cpu().numpy()[ok], st[ok]),
          (wq[ok], wql[ok], wr[ok], wrl[ok], wst[ok]))


def test_adversarial_bytes_unop_slice_lagrange_commit(host, oracle, W):
    import torch
    import util
    rng = np.random.default_rng(92)
    n = 3001
    p, pl, _, _ = util.poly_cases(n, 10, 3, seed=93)
    k = rng.integers(0, 17, n, dtype=np.uint8)
    # shift: 255 cannot fit any row; a shift that exactly fits is fine
    kk = k.copy(); kk[::5] = 255
    out, ol = host.poly_shift(torch.from_numpy(p).cuda(), torch.from_numpy(pl).cuda(), torch.from_numpy(kk).cuda(), 10 + 16)
    out, ol = out.cpu().numpy(), ol.cpu().numpy()
    want, wl = oracle.poly_unop(host.POLY_SHIFT, p, pl, k, 26)
    is_zero = np.array([not p[i, :pl[i]].any() for i in range(n)])
    flagged = (np.arange(n) % 5 == 0) & ~is_zero                     # the zero polynomial shifts to [0] whatever the amount (poly.h:199-216)
    assert (ol[flagged] == 0).all() and not out[flagged].any()
    ps.eq("shift: valid items", (out[~flagged & (np.arange(n) % 5 != 0)], ol[~flagged & (np.arange(n) % 5 != 0)]),
          (want[~flagged & (np.arange(n) % 5 != 0)], wl[~flagged & (np.arange(n) % 5 != 0)]))
    # scale by a byte that is not a field element; a length above the stride
    ks = k.copy(); ks[::4] = 40
    pl2 = pl.copy(); pl2[1::4] = 11
    out, ol = host.poly_scale(torch.from_numpy(p).cuda(), torch.from_numpy(pl2).cuda(), torch.from_numpy(ks).cuda())
    ol = ol.cpu().numpy()
    assert (ol[::4] == 0).all() and (ol[1::4] == 0).all() and (ol[2::4] >= 1).all()
    # slice / lagrange: status 2
    _, _, st = host.poly_slice(torch.from_numpy(p).cuda(), torch.from_numpy(pl2).cuda(), torch.from_numpy(np.zeros(n, np.uint8)).cuda(),
                               torch.from_numpy(np.ones(n, np.uint8)).cuda())
    assert (st.cpu().numpy()[1::4] == 2).all() and (st.cpu().numpy()[2::4] == 0).all()
    xs = np.tile(np.arange(4, dtype=np.uint8), (n, 1)); ys = rng.integers(0, 17, (n, 4), dtype=np.uint8)
    ys[::3, 2] = 17
    _, _, st = host.poly_lagrange(torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda())
    assert (st.cpu().numpy()[::3] == 2).all() and (st.cpu().numpy()[1::3] == 0).all()
    # interpolate_at_h: length 0 for a value byte above 16
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
    vals = rng.integers(0, 17, (n, 4), dtype=np.uint8); vals[::6, 1] = 250
    o, l = pk.interpolate_at_h(torch.from_numpy(vals).cuda())
    assert (l.cpu().numpy()[::6] == 0).all() and not o.cpu().numpy()[::6].any() and (l.cpu().numpy()[1::6] >= 1).all()
    # srs_eval_at_s (table kernel and raw kernel): a coefficient byte that would index past its table row, a length above the stride
    polys = rng.integers(0, 17, (n, 9), dtype=np.uint8); plen = rng.integers(1, 10, n).astype(np.uint8)
    want, wst = pk.srs_eval_at_s(polys, plen)
    polys2, plen2 = polys.copy(), plen.copy()
    polys2[::7, 0] = 200; plen2[3::7] = 77
    for fn in (lambda: pk.srs_eval_at_s(torch.from_numpy(polys2).cuda(), torch.from_numpy(plen2).cuda()),
               lambda: host.srs_eval_at_s_raw(W.generator_srs(9)[0], torch.from_numpy(polys2).cuda(), torch.from_numpy(plen2).cuda())):
        got, st = (t.cpu().numpy() for t in fn())
        assert (st[::7] == 2).all() and (st[3::7] == 2).all() and not got[st == 2].any()
        okm = st != 2
        ps.eq("commit: valid items", (got[okm], st[okm]), (want[okm], wst[okm]))


def test_poly_divide_by_zh(host, oracle, W):
    """pb_poly_divide_zh (the context's Z_H = x^4 - 1, plonk.h:505) == poly_divide with that divisor passed per item: strides 11
    and 22, ragged sizes around the four-items-per-thread groups, zero / short / length-0 numerators, invalid rows."""
    import torch
    rng = np.random.default_rng(33)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    for sn in (11, 22):
        for n in (1, 511, 512, 513, 20011):
            num = rng.integers(0, 17, (n, sn), dtype=np.uint8)
            num[rng.random((n, sn)) < 0.3] = 0
            nl = rng.integers(1, sn + 1, n).astype(np.uint8)      # (a length-0 numerator makes the reference read past a calloc(0): poly.h:133-159)
            num[: n // 10] = 0
            zh = np.tile(np.array([16, 0, 0, 0, 1], np.uint8), (n, 1)); five = np.full(n, 5, np.uint8)
            want = oracle.poly_divide(num, nl, zh, five, sn - 4, 4)
            for path in ("device", "host"):
                got = pk.poly_divide_zh(*( (torch.from_numpy(num).cuda(), torch.from_numpy(nl).cuda()) if path == "device" else (num, nl)))
                got = tuple(np.asarray(g.cpu()) if path == "device" else g for g in got)
                ps.eq(f"divide by Z_H, stride {sn}, n {n}, {path}", got, want)
        num2 = num.copy(); nl2 = nl.copy()
        num2[::9, 0] = 30; nl2[::9] = np.maximum(nl2[::9], 1); nl2[4::9] = sn + 1
        q, ql, r, rl, st = (g.cpu().numpy() for g in pk.poly_divide_zh(torch.from_numpy(num2).cuda(), torch.from_numpy(nl2).cuda()))
        assert (st[::9] == 2).all() and (st[4::9] == 2).all() and not q[st == 2].any() and not r[st == 2].any()
        okm = st != 2
        ps.eq("divide by Z_H: valid items", (q[okm], ql[okm], r[okm], rl[okm], st[okm]), tuple(w[okm] for w in want))
