"""`LatteT2V` — the reference's text-to-video denoiser surface (Vchitect/Latte `models/latte_t2v.py:444-944`), backed by
the same hand-written sm_100a kernels as `Latte` through ONE C-ABI call (`b200_t2v_forward`).

Built for the configuration the reference ships (HF `maxin-cn/Latte-1`): `ada_norm_single`,
LayerNorms without affine, `gelu-approximate`, attention bias, caption projection 4096 -> D, patch 2.  Parameter names
follow the diffusers 0.24.0 state dict of that model, so its `diffusion_pytorch_model.safetensors` loads unchanged.
**Parity is unpinned**: diffusers is not available offline, the CPU truth is `oracle/t2v_oracle.py`'s restatement.
No CPU path.  Not built: attention masks with padding, image joint training (`use_image_num`), LoRA scale, GLIGEN.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .latte import _sincos_1d


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])


class _GELUProj(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FF(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])


class _SpatialBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)
        self.attn1 = _Attn(dim)
        self.attn2 = _Attn(dim)
        self.ff = _FF(dim)


class _TemporalBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)
        self.attn1 = _Attn(dim)
        self.ff = _FF(dim)


class _TimestepEmbedder(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear_1 = nn.Linear(256, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _Emb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedder(dim)


class _AdaLNSingle(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.emb = _Emb(dim)
        self.linear = nn.Linear(dim, 6 * dim)


class _Caption(nn.Module):
    def __init__(self, in_features, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _PatchEmbed(nn.Module):
    def __init__(self, in_channels, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, dim, kernel_size=patch, stride=patch)


class Transformer3DModelOutput(SimpleNamespace):
    pass


class LatteT2V(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=28,
                 patch_size=2, sample_size=64, caption_channels=4096, video_length=16, norm_type="ada_norm_single",
                 activation_fn="gelu-approximate", attention_bias=True, cross_attention_dim=None, **unused):
        super().__init__()
        if norm_type != "ada_norm_single" or activation_fn != "gelu-approximate" or not attention_bias:
            raise NotImplementedError("latte_b200.LatteT2V is built for the Latte-1 configuration only "
                                      "(ada_norm_single, gelu-approximate, attention_bias)")
        D = num_attention_heads * attention_head_dim
        if cross_attention_dim not in (None, D):
            raise NotImplementedError("cross_attention_dim must equal the inner dimension (text is projected to it)")
        self.config = SimpleNamespace(num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                                      in_channels=in_channels, out_channels=out_channels, num_layers=num_layers,
                                      patch_size=patch_size, sample_size=sample_size, caption_channels=caption_channels,
                                      video_length=video_length, norm_type=norm_type)
        self.inner_dim = D
        self.pos_embed = _PatchEmbed(in_channels, D, patch_size)
        self.adaln_single = _AdaLNSingle(D)
        self.caption_projection = _Caption(caption_channels, D)
        self.transformer_blocks = nn.ModuleList([_SpatialBlock(D) for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList([_TemporalBlock(D) for _ in range(num_layers)])
        self.scale_shift_table = nn.Parameter(torch.randn(2, D) / D ** 0.5)
        self.proj_out = nn.Linear(D, patch_size * patch_size * out_channels)
        # non-persistent sin-cos tables, as in the reference (latte_t2v.py:669-671; diffusers PatchEmbed.pos_embed)
        grid = sample_size // patch_size
        scale = max(sample_size // 64, 1)
        coords = np.arange(grid, dtype=np.float32) / scale
        ww, hh = np.meshgrid(coords, coords)
        pos = np.concatenate([_sincos_1d(D // 2, ww), _sincos_1d(D // 2, hh)], axis=1)
        self.register_buffer("pos_table", torch.from_numpy(pos).float(), persistent=False)
        self.register_buffer("temp_pos_embed", torch.from_numpy(_sincos_1d(D, np.arange(video_length))).float().unsqueeze(0),
                             persistent=False)
        self.compute_dtype = torch.float16
        self._packed = None
        self._packed_key = None
        self._workspace = None

    @property
    def dtype(self):
        return self.proj_out.weight.dtype

    @classmethod
    def from_pretrained(cls, path, subfolder=None, video_length=16, torch_dtype=None, **kw):
        """Same call shape as `LatteT2V.from_pretrained(path, subfolder="transformer", video_length=...)`
        (models/__init__.py:41): reads config.json and the diffusers weight file from a local directory."""
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        cfg["video_length"] = video_length
        model = cls(**cfg)
        st = os.path.join(root, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(root, "diffusion_pytorch_model.bin"), map_location="cpu")
        sd = {k: v for k, v in sd.items() if k in model.state_dict()}
        model.load_state_dict(sd, strict=True)
        return model.to(torch_dtype) if torch_dtype is not None else model

    # ---------------------------------------------------------------------------------------------
    def _operand_dtype(self):
        pd = self.proj_out.weight.dtype
        return pd if pd in (torch.float16, torch.bfloat16) else self.compute_dtype

    def repack(self):
        self._packed = None
        self._packed_key = None

    @torch.no_grad()
    def _pack(self):
        ver = sum(p._version for p in self.parameters())
        w0 = self.proj_out.weight
        key = (ver, w0.data_ptr(), w0.device, w0.dtype, self.compute_dtype)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev, od, c = w0.device, self._operand_dtype(), self.config
        D = self.inner_dim
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        h16 = lambda t: t.detach().to(device=dev, dtype=od).contiguous()
        sb, tb = list(self.transformer_blocks), list(self.temporal_transformer_blocks)
        cat = lambda ts: torch.cat([t.detach().reshape(-1) for t in ts])
        stack = lambda ts: torch.stack([t.detach() for t in ts])
        qkv_w = lambda blks, a: stack([torch.cat([getattr(b, a).to_q.weight, getattr(b, a).to_k.weight, getattr(b, a).to_v.weight]) for b in blks])
        qkv_b = lambda blks, a: cat([torch.cat([getattr(b, a).to_q.bias, getattr(b, a).to_k.bias, getattr(b, a).to_v.bias]) for b in blks])
        T = {
            "patch_w": f32(self.pos_embed.proj.weight).reshape(D, -1).contiguous(), "patch_b": f32(self.pos_embed.proj.bias),
            "pos_embed": f32(self.pos_table), "temp_embed": f32(self.temp_pos_embed).reshape(-1, D).contiguous(),
            "t_w0": f32(self.adaln_single.emb.timestep_embedder.linear_1.weight), "t_b0": f32(self.adaln_single.emb.timestep_embedder.linear_1.bias),
            "t_w2": f32(self.adaln_single.emb.timestep_embedder.linear_2.weight), "t_b2": f32(self.adaln_single.emb.timestep_embedder.linear_2.bias),
            "ada_w16": h16(self.adaln_single.linear.weight), "ada_b": f32(self.adaln_single.linear.bias),
            "cap_w1_16": h16(self.caption_projection.linear_1.weight), "cap_b1": f32(self.caption_projection.linear_1.bias),
            "cap_w2_16": h16(self.caption_projection.linear_2.weight), "cap_b2": f32(self.caption_projection.linear_2.bias),
            "tables": f32(torch.stack([t for pair in zip([b.scale_shift_table for b in sb], [b.scale_shift_table for b in tb]) for t in pair])),
            "final_table": f32(self.scale_shift_table),
            "s_qkv_w16": h16(qkv_w(sb, "attn1")), "s_qkv_b": f32(qkv_b(sb, "attn1")),
            "s_out_w16": h16(stack([b.attn1.to_out[0].weight for b in sb])), "s_out_b": f32(cat([b.attn1.to_out[0].bias for b in sb])),
            "c_q_w16": h16(stack([b.attn2.to_q.weight for b in sb])), "c_q_b": f32(cat([b.attn2.to_q.bias for b in sb])),
            "c_kv_w16": h16(torch.cat([torch.cat([b.attn2.to_k.weight.detach(), b.attn2.to_v.weight.detach()]) for b in sb])),
            "c_kv_b": f32(cat([torch.cat([b.attn2.to_k.bias, b.attn2.to_v.bias]) for b in sb])),
            "c_out_w16": h16(stack([b.attn2.to_out[0].weight for b in sb])), "c_out_b": f32(cat([b.attn2.to_out[0].bias for b in sb])),
            "s_fc1_w16": h16(stack([b.ff.net[0].proj.weight for b in sb])), "s_fc1_b": f32(cat([b.ff.net[0].proj.bias for b in sb])),
            "s_fc2_w16": h16(stack([b.ff.net[2].weight for b in sb])), "s_fc2_b": f32(cat([b.ff.net[2].bias for b in sb])),
            "t_qkv_w16": h16(qkv_w(tb, "attn1")), "t_qkv_b": f32(qkv_b(tb, "attn1")),
            "t_out_w16": h16(stack([b.attn1.to_out[0].weight for b in tb])), "t_out_b": f32(cat([b.attn1.to_out[0].bias for b in tb])),
            "t_fc1_w16": h16(stack([b.ff.net[0].proj.weight for b in tb])), "t_fc1_b": f32(cat([b.ff.net[0].proj.bias for b in tb])),
            "t_fc2_w16": h16(stack([b.ff.net[2].weight for b in tb])), "t_fc2_b": f32(cat([b.ff.net[2].bias for b in tb])),
            "final_w": f32(self.proj_out.weight), "final_b": f32(self.proj_out.bias),
        }
        w = _lib.T2VWeights()
        for name in _lib.T2V_WEIGHT_FIELDS:
            setattr(w, name, T[name].data_ptr())
        shape = _lib.T2VShape(layers=c.num_layers, hidden=D, heads=c.num_attention_heads, mlp_hidden=4 * D, patch=c.patch_size,
                              in_channels=c.in_channels, out_channels=c.out_channels, input_size=c.sample_size,
                              frames=c.video_length, caption_channels=c.caption_channels,
                              dtype=_lib.BF16 if od == torch.bfloat16 else _lib.FP16)
        self._packed, self._packed_key = (shape, w, T), key
        return self._packed

    def forward(self, hidden_states, timestep=None, encoder_hidden_states=None, added_cond_kwargs=None,
                class_labels=None, cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None,
                use_image_num=0, enable_temporal_attentions=True, return_dict=True):
        """hidden_states (B, C, F, H, W), timestep (B,), encoder_hidden_states (B, L<=128, caption_channels)
        -> (B, out_channels, F, H, W)  (latte_t2v.py:677-941)."""
        if not hidden_states.is_cuda:
            raise RuntimeError("latte_b200.LatteT2V runs on CUDA (sm_100a) only; there is no CPU fallback")
        if use_image_num != 0 or attention_mask is not None or cross_attention_kwargs:
            raise NotImplementedError("use_image_num / attention_mask / cross_attention_kwargs are not built")
        if encoder_attention_mask is not None and not bool(encoder_attention_mask.bool().all()):
            raise NotImplementedError("padded text (encoder_attention_mask with zeros) is not built; truncate the prompt "
                                      "embeddings to the true token count as pipeline_latte.py:118-124 does for batch 1")
        c = self.config
        B = hidden_states.shape[0]
        if tuple(hidden_states.shape[1:]) != (c.in_channels, c.video_length, c.sample_size, c.sample_size):
            raise ValueError(f"hidden_states must be (B, {c.in_channels}, {c.video_length}, {c.sample_size}, {c.sample_size})")
        dev = hidden_states.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            shape, w, _ = self._pack()
            x = hidden_states.detach().to(torch.float32).contiguous()
            t = timestep.detach().to(device=dev, dtype=torch.int64).reshape(-1).expand(B).contiguous()
            text = encoder_hidden_states.detach().to(device=dev, dtype=torch.float32).contiguous()
            if text.dim() != 3 or text.shape[0] != B or text.shape[2] != c.caption_channels:
                raise ValueError("encoder_hidden_states must be (B, L, caption_channels)")
            L = text.shape[1]
            out = torch.empty(B, c.out_channels, c.video_length, c.sample_size, c.sample_size, dtype=torch.float32, device=dev)
            need = lib.b200_t2v_workspace_bytes(C.byref(shape), B, L)
            if need == 0:
                raise RuntimeError("latte_b200: unsupported T2V configuration: " + _lib.last_error())
            ws = self._workspace
            if ws is None or ws.numel() < need + 1024 or ws.device != dev:
                ws = self._workspace = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            rc = lib.b200_t2v_forward(C.byref(shape), C.byref(w), x.data_ptr(), t.data_ptr(), text.data_ptr(), B, L,
                                      int(bool(enable_temporal_attentions)), out.data_ptr(), base, need,
                                      torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "b200_t2v_forward")
        pd = self.dtype
        out = out if pd == torch.float32 else out.to(pd)
        return Transformer3DModelOutput(sample=out) if return_dict else (out,)
