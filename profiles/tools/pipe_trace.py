"""Time the host-pointer pipelines alone (no bench around them): python profiles/tools/pipe_trace.py [struct|compact|packed|packed3] [reps]
With PB_PIPE_TRACE=1 the library prints, per chunk, when its inputs landed / its kernels finished / its outputs were copied out.
PB_PIPE_CHUNK=<items> overrides the chunk size (read once per process)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from plonk_c_b200 import host, wire, workload as W
mode = sys.argv[1] if len(sys.argv) > 1 else "struct"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
n = 1 << 21
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
batch = W.make_batch(5, 0, n)
wit, rnd, chal, u = [pin(x) for x in batch]
packed = pin(wire.pack_inputs(*batch))
out = [pin(np.empty((n, 34), np.uint8)), pin(np.empty(n, np.uint8)), pin(np.empty(n, np.uint8))]
pout = [pin(np.empty((n, 22), np.uint8)), pin(np.empty(n, np.uint8))]
packed3 = pin(wire.pack_inputs3(*batch))
pout3 = [pin(np.empty((n, 12), np.uint8)), pin(np.empty(n, np.uint8))]
fn = {"struct": lambda: pk.prove_verify_into(wit, rnd, chal, u, *out),
      "compact": lambda: pk.prove_verify_compact_into(wit, rnd, chal, u, *out),
      "packed": lambda: pk.prove_verify_packed_into(packed, *pout),
      "packed3": lambda: pk.prove_verify_packed3_into(packed3, *pout3)}[mode]
ts = []
for i in range(reps + 2):
    t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
print(f"{mode}: calls " + " ".join(f"{t:.3f}" for t in ts) + f" ms; best {min(ts):.3f} ms = {n / min(ts) / 1e6:.3f} G proofs/s", file=sys.stderr)
