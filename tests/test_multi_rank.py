"""world_size-2 test of the N > 1 path on CPU (gloo): contiguous sharding of the item range, per-rank work on the
rank's own slice of the counter-based stream, and the one collective of the path -- all-reduce(sum) of the
counters, all-reduce(max) of the elapsed time.  The per-rank compute is the CPU oracle
here (test infrastructure); on GPUs it is the CUDA library, and bench.py runs this same reduction over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL = 6001       # deliberately not divisible by the world size


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle._binding import load_port
    from plonk_c_b200 import shard, workload as W
    lo, cnt = shard.shard_range(N_TOTAL, rank, world)
    wit, rnd, chal, u = W.make_batch(99, lo, cnt, "U17")
    oracle = load_port()
    g1s, g2 = W.generator_srs(9)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    verdict = np.where(status == 0, verdict, 0xFF).astype(np.uint8)
    counts = torch.from_numpy(shard.tally_host(proofs, status, verdict))
    total, elapsed = shard.reduce_counters(counts, elapsed_ms=10.0 * (rank + 1))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([total, [elapsed, lo, cnt]]))
    dist.destroy_process_group()


def test_sharded_prove_verify_reduction_world2(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npy") for k in range(world)]
    # both ranks hold the same reduced counters and the max elapsed time
    assert np.array_equal(r[0][:19], r[1][:19]) and r[0][18] == 20.0
    # shards are contiguous, disjoint and cover [0, N_TOTAL)
    assert r[0][19] == 0 and r[0][19] + r[0][20] == r[1][19] and r[1][19] + r[1][20] == N_TOTAL
    # the reduced counters equal a single-process run over the whole range
    sys.path.insert(0, ROOT)
    from oracle._binding import load_port
    from plonk_c_b200 import shard, workload as W
    wit, rnd, chal, u = W.make_batch(99, 0, N_TOTAL, "U17")
    g1s, g2 = W.generator_srs(9)
    o = load_port()
    proofs, status = o.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = o.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    verdict = np.where(status == 0, verdict, 0xFF).astype(np.uint8)
    want = shard.tally_host(proofs, status, verdict)
    assert np.array_equal(r[0][:18].astype(np.int64), want)
    assert want[:16].sum() == N_TOTAL and want[0] > 0 and want[8] > 0


@pytest.mark.parametrize("n,world", [(16, 8), (17, 8), (1 << 24, 8), (5, 8), (0, 4), (1000003, 3)])
def test_shard_range_partitions(n, world):
    sys.path.insert(0, ROOT)
    from plonk_c_b200 import shard
    pos = 0
    for r in range(world):
        lo, cnt = shard.shard_range(n, r, world)
        assert lo == pos and cnt >= 0
        pos += cnt
    assert pos == n
    sizes = [shard.shard_range(n, r, world)[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
