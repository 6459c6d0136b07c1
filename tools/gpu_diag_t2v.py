"""Bring-up diagnostics for the LatteT2V pieces (cross attention, long attention, whole forward)."""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latte_b200 import ops, LatteT2V
from oracle import t2v_oracle as T
dev = torch.device("cuda:0")

def stat(name, got, ref, tol=4e-3):
    err = (got.float() - ref.float()).abs()
    bad = err > tol + tol * ref.float().abs()
    print(f"[{'OK ' if not bad.any() else 'BAD'}] {name}: maxabs {err.max().item():.3e} mean {err.mean().item():.3e} ref absmax {ref.abs().max().item():.3f} bad {bad.float().mean().item()*100:.2f}% nan {torch.isnan(got.float()).sum().item()}", flush=True)
    return bad

def cross():
    g = torch.Generator().manual_seed(0)
    for (b, rows, L, h, hd) in [(1, 128, 128, 1, 64), (1, 128, 20, 1, 64), (2, 256, 120, 4, 72)]:
        D = h * hd
        q = torch.randn(b * rows, D, generator=g).to(dev).half()
        kv = torch.randn(b * L, 2 * D, generator=g).to(dev).half()
        for label, qq, kk in [("V=1", q, torch.cat([kv[:, :D], torch.ones_like(kv[:, D:])], 1).contiguous()), ("Q=0", torch.zeros_like(q), kv), ("rand", q, kv)]:
            try:
                out = ops.cross_attention(qq, kk, b, rows, L, h)
            except Exception as e:
                print("ERR", label, e); continue
            qf = qq.float().reshape(b, rows, h, hd).transpose(1, 2)
            kf = kk.float()[:, :D].reshape(b, L, h, hd).transpose(1, 2)
            vf = kk.float()[:, D:].reshape(b, L, h, hd).transpose(1, 2)
            ref = (torch.softmax(qf @ kf.transpose(-1, -2) * hd ** -0.5, dim=-1) @ vf).transpose(1, 2).reshape(b * rows, D)
            bad = stat(f"cross b{b} rows{rows} L{L} h{h} hd{hd} {label}", out, ref)
            if bad.any():
                print("   bad per column:", bad.float().mean(0).cpu().numpy().round(2)[:16], " per row[:16]:", bad.float().mean(1).cpu().numpy().round(2)[:16])

def model():
    for kw, batch, L, temporal in [
        (dict(num_attention_heads=2, attention_head_dim=64, num_layers=1, sample_size=16, video_length=8, caption_channels=256), 2, 16, False),
        (dict(num_attention_heads=2, attention_head_dim=64, num_layers=2, sample_size=16, video_length=8, caption_channels=256), 2, 20, True),
        (dict(num_attention_heads=8, attention_head_dim=72, num_layers=2, sample_size=32, video_length=16, caption_channels=512), 2, 120, True),
        (dict(num_attention_heads=8, attention_head_dim=72, num_layers=1, sample_size=64, video_length=4, caption_channels=256), 1, 33, True)]:
        try:
            cfg = T.T2VConfig(**kw); sd = T.make_weights(cfg, 3); x, t, text = T.make_inputs(cfg, batch, L, 4)
            net = LatteT2V(**kw); net.load_state_dict(sd); net = net.cuda().eval()
            with torch.no_grad():
                out = net(x.cuda(), t.cuda(), encoder_hidden_states=text.cuda(), enable_temporal_attentions=temporal, return_dict=False)[0]
            torch.cuda.synchronize()
            stat(f"t2v {kw} b{batch} L{L} temporal={temporal}", out.cpu(), T.t2v_forward(sd, cfg, x, t, text, enable_temporal=temporal), 1e-2)
        except Exception as e:
            print("ERR", kw, repr(e)[:300])

if __name__ == "__main__":
    {"cross": cross, "model": model}[sys.argv[1]]()
