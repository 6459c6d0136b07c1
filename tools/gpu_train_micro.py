"""Microbenchmark of the training-step reduction kernels at Latte-XL/2, local batch 5 (T = 20480 rows).  Knobs come from the
environment (read once per process), so the driver below re-runs this file per configuration."""
import os, subprocess, sys
import torch
sys.path.insert(0, ".")

CONFIGS = [
    {},
    {"B200_TRAIN_DBG": "1"},
    {"B200_TRAIN_RS_GELU": "64", "B200_TRAIN_RS_GATE": "64", "B200_TRAIN_RS_COLSUM": "128", "B200_LNB_WARPS": "8", "B200_LNB_RPW": "8"},
    {"B200_TRAIN_RS_GELU": "128", "B200_TRAIN_RS_GATE": "128", "B200_TRAIN_RS_COLSUM": "256", "B200_LNB_WARPS": "8", "B200_LNB_RPW": "4"},
    {"B200_TRAIN_RS_GELU": "256", "B200_TRAIN_RS_GATE": "256", "B200_TRAIN_RS_COLSUM": "512", "B200_LNB_WARPS": "4", "B200_LNB_RPW": "16"},
    {"B200_TRAIN_RS_GELU": "16", "B200_TRAIN_RS_GATE": "16", "B200_TRAIN_RS_COLSUM": "32", "B200_LNB_WARPS": "2", "B200_LNB_RPW": "8"},
]


def worker():
    from latte_b200.train_ops import NativeOps
    dev = torch.device("cuda:0")
    ops = NativeOps(torch.bfloat16)
    B, rpb, D = 5, 4096, 1152
    T = B * rpb
    g = torch.Generator().manual_seed(0)
    dx = torch.randn(T, D, generator=g).to(dev)
    x = torch.randn(T, D, generator=g).to(dev)
    m16 = torch.randn(T, D, generator=g).to(dev).bfloat16()
    u = torch.randn(T, 4 * D, generator=g).to(dev).bfloat16()
    da = torch.randn(T, 4 * D, generator=g).to(dev).bfloat16()
    dqkv = torch.randn(T, 3 * D, generator=g).to(dev).bfloat16()
    mod = torch.randn(B, 6 * D, generator=g).to(dev)
    dmod = torch.zeros(B, 6 * D, device=dev)
    db = torch.zeros(4 * D, device=dev)
    qkv = torch.randn(T, 3 * D, generator=g).to(dev).bfloat16()
    o = ops.attention(qkv, B, 16, 256, 16, False)
    do = torch.randn(T, D, generator=g).to(dev).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def bench(fn, n=10):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / n * 1000

    res = {
        "gate_bwd": bench(lambda: ops.gate_bwd(dx, m16, mod[:, 2 * D:3 * D], rpb, dmod[:, 2 * D:3 * D], db[:D])),
        "gelu_bwd": bench(lambda: ops.gelu_bwd(da, u, db)),
        "ln_bwd": bench(lambda: ops.ln_modulate_bwd(m16, x, mod[:, :D], mod[:, D:2 * D], rpb, dx, dmod[:, :D], dmod[:, D:2 * D])),
        "colsum": bench(lambda: ops.colsum(dqkv, db[:3 * D])),
        "attn_bwd_sp": bench(lambda: ops.attention_bwd(qkv, o, do, B, 16, 256, 16, False)),
        "attn_bwd_tmp": bench(lambda: ops.attention_bwd(qkv, o, do, B, 16, 256, 16, True)),
        "gelu": bench(lambda: ops.gelu(u)),
    }
    print({k: v for k, v in os.environ.items() if k.startswith("B200_")}, {k: round(v, 1) for k, v in res.items()}, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
    else:
        for cfg in CONFIGS:
            env = dict(os.environ); env.update(cfg)
            subprocess.run([sys.executable, __file__, "worker"], env=env, timeout=300)
