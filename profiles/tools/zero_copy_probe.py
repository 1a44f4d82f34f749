"""Can the kernels stream their batch over PCIe themselves?  Runs pb_plonk_prove_dev with pinned HOST buffers passed as
device pointers (unified addressing: the SMs read / write host memory directly) and compares with device-resident buffers.
usage: python profiles/tools/zero_copy_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from plonk_c_b200 import host, workload as W  # noqa: E402

n = 1 << 21
lib = host.lib()
pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
wit, rnd, chal, u = W.make_batch(2025, 0, n, "U17")
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)


def P(t):
    return C.c_void_p(t.data_ptr())


def bufs(where_in, where_out):
    mk_in = (lambda x: torch.from_numpy(x).pin_memory()) if where_in == "host" else (lambda x: torch.from_numpy(x).cuda())
    if where_out == "host":
        mk_out = lambda *s: torch.empty(s, dtype=torch.uint8).pin_memory()
    else:
        mk_out = lambda *s: torch.empty(s, dtype=torch.uint8, device="cuda")
    return [mk_in(x) for x in (wit, rnd, chal)], [mk_out(n, 34), mk_out(n)]


ref = None
for where_in, where_out in (("device", "device"), ("host", "device"), ("device", "host"), ("host", "host")):
    ins, outs = bufs(where_in, where_out)
    times = []
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        host._check(lib.pb_plonk_prove_dev(pk._h, P(ins[0]), P(ins[1]), P(ins[2]), P(outs[0]), P(outs[1]), C.c_size_t(n), sp))
        b.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    cs = int(outs[0].cpu().to(torch.int64).sum())
    ref = ref or cs
    t = min(times[1:])
    print(f"in={where_in:6s} out={where_out:6s}  prove {t*1e3:8.1f} us   in {27*n/t/1e6:6.1f} GB/s  out {35*n/t/1e6:6.1f} GB/s  checksum ok {cs == ref}")
