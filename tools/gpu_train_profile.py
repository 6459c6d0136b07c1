"""Per-op device time of one Latte-XL/2 training step (local batch 5): wraps every NativeOps method with CUDA events."""
import collections, json, sys
import torch
sys.path.insert(0, ".")
from latte_b200 import Latte_models, training, train_ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Latte_models["Latte-XL/2"](input_size=32, num_classes=101, num_frames=16, learn_sigma=True, extras=2).to(dev)
with torch.no_grad():
    for p in m.parameters():
        if p.requires_grad and float(p.abs().max()) == 0.0:
            p.normal_(0, 0.02)
ops = train_ops.NativeOps(torch.bfloat16)
events = []
def wrap(name, fn):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        tag = name
        if name in ("linear", "linear_accum"):
            M = a[0].shape[0]; N = a[1].shape[0] if name == "linear" else a[0].shape[1]; K = a[1].shape[1] if name == "linear" else a[1].shape[1]
            tag = f"{name} {M}x{N}x{K}"
        if name == "dgrad":
            tag = f"dgrad {a[0].shape[0]}x{a[0].shape[1]}x{a[1].shape[1]}"
        if name == "wgrad":
            tag = f"wgrad {a[1].shape[0]}x{a[1].shape[1]}x{a[2].shape[1]}"
        if name == "attention_bwd":
            tag = "attention_bwd " + ("temporal" if a[-1] else "spatial")
        events.append((tag, e0, e1))
        return r
    return inner
for n in ("ln_modulate", "linear", "linear_accum", "attention", "gate_residual", "gelu", "gate_bwd", "gelu_bwd", "ln_modulate_bwd",
          "attention_bwd", "wgrad", "dgrad", "cast_into", "linear_gelu_both", "colsum", "transpose", "cast", "to_operand", "ada_outer", "ada_dsc"):
    setattr(ops, n, wrap(n, getattr(ops, n)))
x = torch.randn(B, 16, 4, 32, 32, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
y = torch.randint(0, 101, (B,), device=dev)
dout = torch.randn(B, 16, 8, 32, 32, device=dev)
eng = training.TrainEngine(m, ops, torch.bfloat16)
res = {}
for it in range(2):
    events.clear()
    torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        marks[0].record(); eng.prepare(); marks[1].record()
        c = training.conditioning(m, t, y)
        out = eng.forward(x, c); marks[2].record()
        G, dc = eng.backward(dout); marks[3].record()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for tag, e0, e1 in events:
    d = agg.setdefault(tag, [0, 0.0]); d[0] += 1; d[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"prepare {marks[0].elapsed_time(marks[1]):.2f} ms, forward {marks[1].elapsed_time(marks[2]):.2f} ms, backward {marks[2].elapsed_time(marks[3]):.2f} ms; sum of op times {tot:.2f} ms")
for tag, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms  {n:5d} x  {tag}")
