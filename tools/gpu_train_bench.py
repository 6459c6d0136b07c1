"""Config-5 timing (not a bench line): Latte-XL/2 fwd + bwd at local batch 5 through model(x,t,y) / loss.backward()."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from latte_b200 import Latte_models
from latte_b200.diffusion import create_diffusion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Latte_models["Latte-XL/2"](input_size=32, num_classes=101, num_frames=16, learn_sigma=True, extras=2).to(dev)
with torch.no_grad():
    for p in m.parameters():
        if p.requires_grad and float(p.abs().max()) == 0.0:
            p.normal_(0, 0.02)
m.train()
d = create_diffusion(timestep_respacing="")
x = torch.randn(B, 16, 4, 32, 32, device=dev)
y = torch.randint(0, 101, (B,), device=dev)
def step():
    t = torch.randint(0, 1000, (B,), device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = d.training_losses(m, x, t, dict(y=y))["loss"].mean()
    m.zero_grad(set_to_none=True)
    loss.backward()
    return loss
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
t0 = time.time(); e0.record()
for _ in range(n):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"workload": f"Latte-XL/2 train fwd+bwd, local batch {B}, bf16 operands", "ms_per_step": ms, "steps_per_s": 1000 / ms,
                  "wall_ms": (time.time() - t0) * 1000 / n, "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                  "tflops_fwd_bwd": 3 * 3.7256 * B / ms}))
