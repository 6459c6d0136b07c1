"""AutoencoderKL.encode of train.py's per-step batch: local batch 5 x 16 frames of 3x256x256."""
import json, sys
import torch
sys.path.insert(0, ".")
from latte_b200 import AutoencoderKL
from oracle import vae_oracle as V
cfg = V.VaeConfig()
vae = AutoencoderKL()
vae.load_state_dict(V.make_weights(cfg, 1), strict=True)
vae = vae.cuda().eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
x = torch.rand(n, 3, 256, 256, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(2):
        z = vae.encode(x).latent_dist.sample().mul_(0.18215)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        z = vae.encode(x).latent_dist.sample().mul_(0.18215)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"workload": f"AutoencoderKL.encode, {n} frames 3x256x256 -> 4x32x32, fp16", "ms": ms, "ms_per_frame": ms / n,
                  "tflops_achieved": n * 0.27 / (ms * 1e-3) / 1e3 * 1e3 / 1e3 * 1e3 if False else n * 0.27 / (ms * 1e-3), "latent_shape": list(z.shape)}))
