"""The .pbatch on-disk format (plonk.c_b200/wire.py): round trip, and the oracle's results stored next to the inputs."""
import numpy as np
import pytest

from plonk_c_b200 import wire


def test_round_trip_with_oracle_results(tmp_path, oracle, W):
    n = 3001
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(8, 0, n)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    path = tmp_path / "batch.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u, proofs, status, verdict)
    assert path.stat().st_size == 28 + 44 + 30 + 4 + n * (12 + 9 + 5 + 1 + 34 + 1 + 1)
    b = wire.read_batch(path)
    for k, v in dict(circuit=W.PLONK_TEST_CIRCUIT, srs_g1s=g1s, srs_g2=g2, witness=wit, rand=rnd, chal=chal, u=u,
                     proofs=proofs, status=status, verdict=verdict).items():
        assert np.array_equal(b[k], v), k
    # inputs only
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
    b = wire.read_batch(path)
    assert b["flags"] == 0 and "proofs" not in b and b["n"] == n


def test_rejects_garbage(tmp_path):
    p = tmp_path / "x.pbatch"
    p.write_bytes(b"not a batch file at all, definitely not" * 3)
    with pytest.raises(ValueError):
        wire.read_batch(p)


@pytest.mark.gpu
def test_file_to_gpu_to_file(tmp_path, host, oracle, W):
    """A batch file is proved and verified on the GPU straight from its memory-mapped arrays; results equal the oracle's."""
    n = 20011
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(9, 0, n)
    path = tmp_path / "in.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
    b = wire.read_batch(path)
    pk = host.Plonk(b["circuit"], b["srs_g1s"], b["srs_g2"])
    proofs, status, verdict = pk.prove_verify(np.array(b["witness"]), np.array(b["rand"]), np.array(b["chal"]), np.array(b["u"]))
    rp, rs = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    rv, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, rp, chal, u)
    rv = np.where(rs == 0, rv, 0xFF).astype(np.uint8)
    assert np.array_equal(proofs, rp) and np.array_equal(status, rs) and np.array_equal(verdict, rv)
