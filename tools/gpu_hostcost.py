"""Host-side cost of one denoising step through the public module call, eager launch sequence vs CUDA-graph replay
(VERDICT r01 #2): enqueue time per step with no synchronisation (what a sampling loop pays), and the synchronous
latency of one step (what bench.py's e2e leg pays).  Under torchrun every rank prints its own line (N = 8: host contention)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latte_b200 import Latte  # noqa: E402
from oracle import latte_oracle as O  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = O.make_config("Latte-XL/2")
sd = O.make_weights(cfg, 0)
x, t, y = O.make_inputs(cfg, 2, 123)
net = Latte(input_size=cfg.input_size, hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads,
            num_frames=cfg.num_frames, num_classes=cfg.num_classes, learn_sigma=True, extras=2)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
xd, td, yd = x.to(dev), t.to(dev), y.to(dev)
n = 40
with torch.no_grad():
    for graphs in (False, True):
        net.use_cuda_graphs = graphs
        for _ in range(4):
            net.forward_with_cfg(xd, td, y=yd, cfg_scale=7.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            net.forward_with_cfg(xd, td, y=yd, cfg_scale=7.0)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        lat = []
        for _ in range(10):
            torch.cuda.synchronize()
            a = time.perf_counter()
            net.forward_with_cfg(xd, td, y=yd, cfg_scale=7.0)
            torch.cuda.synchronize()
            lat.append(time.perf_counter() - a)
        lat.sort()
        print(f"rank {rank} {'graph' if graphs else 'eager'}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, "
              f"loop {1e3 * (t2 - t0) / n:.3f} ms/step, synchronous step latency median {1e3 * lat[len(lat) // 2]:.3f} ms", flush=True)
