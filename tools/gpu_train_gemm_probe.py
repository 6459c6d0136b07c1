"""The backward GEMMs of one Latte-XL/2 block at local batch 5 (T = 20480 tokens), for `ncu --set full` and for event timing:
dgrad (W read in place, MN-major W operand) and wgrad (both operands MN-major, fp32 accumulate) of fc1 / fc2 / qkv / proj."""
import sys
import torch
sys.path.insert(0, ".")
from latte_b200.train_ops import NativeOps

dev = torch.device("cuda:0")
ops = NativeOps(torch.bfloat16)
T, D = 20480, 1152
g = torch.Generator().manual_seed(0)
shapes = {"fc1": (4 * D, D), "fc2": (D, 4 * D), "qkv": (3 * D, D), "proj": (D, D)}
res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (n_out, n_in) in shapes.items():
    dy = torch.randn(T, n_out, generator=g).to(dev).bfloat16()
    x = torch.randn(T, n_in, generator=g).to(dev).bfloat16()
    w = (torch.randn(n_out, n_in, generator=g) / n_in ** 0.5).to(dev).bfloat16()
    gw = torch.zeros(n_out, n_in, device=dev)
    for kind, fn in (("dgrad", lambda: ops.dgrad(dy, w)), ("wgrad", lambda: ops.wgrad(gw, dy, x))):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / 5 * 1000
        res[f"{kind} {name}"] = (us, 2.0 * T * n_out * n_in / (us * 1e-6) / 1e12)
for k, (us, tf) in res.items():
    print(f"{k:12s} {us:8.1f} us  {tf:7.1f} TFLOP/s")
