import sys, numpy as np, torch, time
sys.path.insert(0, ".")
from plonk_c_b200 import host, workload as W
pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
n = 1 << 21
wit, rnd, chal, u = [torch.from_numpy(x).pin_memory().numpy() for x in W.make_batch(5, 0, n)]
out = [torch.empty((n, 34), dtype=torch.uint8).pin_memory().numpy(), torch.empty(n, dtype=torch.uint8).pin_memory().numpy(), torch.empty(n, dtype=torch.uint8).pin_memory().numpy()]
for i in range(3):
    t = time.perf_counter(); pk.prove_verify_into(wit, rnd, chal, u, *out); print("call %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), file=sys.stderr)
