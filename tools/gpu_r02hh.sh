#!/bin/bash
cd "$(dirname "$0")/.."
export LATTE_B200_NO_BUILD=1
cat > /tmp/grl.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from latte_b200.train_ops import NativeOps
dev = torch.device("cuda:0"); ops = NativeOps(torch.bfloat16)
B, rpb, D = 5, 4096, 1152; T = B * rpb
g = torch.Generator().manual_seed(0)
x = torch.randn(T, D, generator=g).to(dev); m16 = torch.randn(T, D, generator=g).to(dev).bfloat16(); mod = torch.randn(B, 6 * D, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(fn, n=10):
    fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1000
print("gate_residual_ln", round(bench(lambda: ops.gate_residual_ln(x, m16, mod[:, 2*D:3*D], mod[:, :D], mod[:, D:2*D], rpb)), 1), "us;  separate:",
      round(bench(lambda: ops.gate_residual(x, m16, mod[:, 2*D:3*D], rpb)), 1), "+", round(bench(lambda: ops.ln_modulate(x, mod[:, :D], mod[:, D:2*D], rpb)), 1), "us")
PY
echo "--- capped (96 regs)"; python /tmp/grl.py
echo "--- uncapped"; B200_GRL_MINB=1 python /tmp/grl.py
timeout 300 python -m pytest tests/test_gpu_train.py -q -k "elementwise_forward or reference_gradients or xl_head" 2>&1 | tail -2
timeout 300 python tools/gpu_train_bench.py 5 2>/dev/null
