"""CPU oracle for AutoencoderKL.decode / .encode — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  **Parity unpinned.**

The reference calls `vae.decode(z).sample` (sample/sample.py:114, sample/sample_ddp.py:167,
sample/pipeline_latte.py:758,771) on diffusers' `AutoencoderKL` (diffusers==0.24.0, environment.yml:13), which is
neither vendored in /root/reference nor installed in this image.  This file restates the published 0.24.0 decoder
(SURVEY.md App. C.4): post_quant_conv 1x1 -> Decoder[conv_in -> UNetMidBlock2D(ResnetBlock2D, single-head Attention over
h*w with GroupNorm, ResnetBlock2D) -> UpDecoderBlock2D x n (layers_per_block+1 ResnetBlock2D each, nearest-2x
Upsample2D + conv on all but the last) -> GroupNorm -> SiLU -> conv_out].  State-dict key names follow diffusers.
No reference-generated golden exists, so parity claims resting on this file are capped at "partial".
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VaeConfig:
    latent_channels: int = 4
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @property
    def up_channels(self):
        return tuple(reversed(self.block_out_channels))


def _resnet_spec(prefix, cin, cout):
    s = [(f"{prefix}.norm1.weight", (cin,)), (f"{prefix}.norm1.bias", (cin,)),
         (f"{prefix}.conv1.weight", (cout, cin, 3, 3)), (f"{prefix}.conv1.bias", (cout,)),
         (f"{prefix}.norm2.weight", (cout,)), (f"{prefix}.norm2.bias", (cout,)),
         (f"{prefix}.conv2.weight", (cout, cout, 3, 3)), (f"{prefix}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{prefix}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{prefix}.conv_shortcut.bias", (cout,))]
    return s


def state_dict_spec(cfg: VaeConfig):
    up = cfg.up_channels
    C0 = up[0]
    spec = [("post_quant_conv.weight", (cfg.latent_channels, cfg.latent_channels, 1, 1)), ("post_quant_conv.bias", (cfg.latent_channels,)),
            ("decoder.conv_in.weight", (C0, cfg.latent_channels, 3, 3)), ("decoder.conv_in.bias", (C0,))]
    spec += _resnet_spec("decoder.mid_block.resnets.0", C0, C0)
    a = "decoder.mid_block.attentions.0"
    spec += [(f"{a}.group_norm.weight", (C0,)), (f"{a}.group_norm.bias", (C0,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        spec += [(f"{a}.{n}.weight", (C0, C0)), (f"{a}.{n}.bias", (C0,))]
    spec += _resnet_spec("decoder.mid_block.resnets.1", C0, C0)
    cin = C0
    for b, co in enumerate(up):
        for r in range(cfg.layers_per_block + 1):
            spec += _resnet_spec(f"decoder.up_blocks.{b}.resnets.{r}", cin if r == 0 else co, co)
        if b + 1 < len(up):
            spec += [(f"decoder.up_blocks.{b}.upsamplers.0.conv.weight", (co, co, 3, 3)), (f"decoder.up_blocks.{b}.upsamplers.0.conv.bias", (co,))]
        cin = co
    spec += [("decoder.conv_norm_out.weight", (up[-1],)), ("decoder.conv_norm_out.bias", (up[-1],)),
             ("decoder.conv_out.weight", (cfg.out_channels, up[-1], 3, 3)), ("decoder.conv_out.bias", (cfg.out_channels,))]
    return spec


def encoder_state_dict_spec(cfg: VaeConfig):
    """diffusers 0.24.0 `Encoder` + `quant_conv` keys: conv_in, DownEncoderBlock2D x n (layers_per_block
    resnets, Downsample2D conv on all but the last), UNetMidBlock2D, conv_norm_out, conv_out (2 * latent channels)."""
    ch = cfg.block_out_channels
    spec = [("encoder.conv_in.weight", (ch[0], cfg.in_channels, 3, 3)), ("encoder.conv_in.bias", (ch[0],))]
    cin = ch[0]
    for b, co in enumerate(ch):
        for r in range(cfg.layers_per_block):
            spec += _resnet_spec(f"encoder.down_blocks.{b}.resnets.{r}", cin if r == 0 else co, co)
        if b + 1 < len(ch):
            spec += [(f"encoder.down_blocks.{b}.downsamplers.0.conv.weight", (co, co, 3, 3)), (f"encoder.down_blocks.{b}.downsamplers.0.conv.bias", (co,))]
        cin = co
    C0 = ch[-1]
    spec += _resnet_spec("encoder.mid_block.resnets.0", C0, C0)
    a = "encoder.mid_block.attentions.0"
    spec += [(f"{a}.group_norm.weight", (C0,)), (f"{a}.group_norm.bias", (C0,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        spec += [(f"{a}.{n}.weight", (C0, C0)), (f"{a}.{n}.bias", (C0,))]
    spec += _resnet_spec("encoder.mid_block.resnets.1", C0, C0)
    M = 2 * cfg.latent_channels
    spec += [("encoder.conv_norm_out.weight", (C0,)), ("encoder.conv_norm_out.bias", (C0,)),
             ("encoder.conv_out.weight", (M, C0, 3, 3)), ("encoder.conv_out.bias", (M,)),
             ("quant_conv.weight", (M, M, 1, 1)), ("quant_conv.bias", (M,))]
    return spec


def make_weights(cfg: VaeConfig, seed: int = 0):
    """Decoder keys first (drawn exactly as before the encoder was added, so existing expectations keep their numbers), then the
    encoder + quant_conv keys from a second generator."""
    g = torch.Generator().manual_seed(seed)
    g_enc = torch.Generator().manual_seed(seed + 7919)
    sd = {}
    for name, shape in state_dict_spec(cfg) + encoder_state_dict_spec(cfg):
        if name.startswith(("encoder.", "quant_conv.")):
            g = g_enc
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("norm.weight") or name.endswith("norm_out.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = t.contiguous()
    return sd


def resnet(sd, p, x, groups):
    h = F.conv2d(F.silu(F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-6)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps=1e-6)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def mid_attention(sd, p, x, groups):
    n, c, h, w = x.shape
    r = x
    t = F.group_norm(x.reshape(n, c, h * w), groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], eps=1e-6).transpose(1, 2)
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1)
    o = F.linear(a @ v, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(n, c, h, w) + r


def vae_encode(sd, cfg: VaeConfig, x):
    """x (n, 3, H, W) -> moments (n, 2 * latent_channels, H/8, W/8) = quant_conv(Encoder(x)): mean | logvar of the
    DiagonalGaussianDistribution that train.py:206-211 samples.  Downsample2D(padding=0) = F.pad (0, 1, 0, 1) + Conv2d stride 2."""
    G = cfg.norm_num_groups
    ch = cfg.block_out_channels
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for b in range(len(ch)):
        for r in range(cfg.layers_per_block):
            h = resnet(sd, f"encoder.down_blocks.{b}.resnets.{r}", h, G)
        if b + 1 < len(ch):
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0.0)
            h = F.conv2d(h, sd[f"encoder.down_blocks.{b}.downsamplers.0.conv.weight"], sd[f"encoder.down_blocks.{b}.downsamplers.0.conv.bias"], stride=2)
    h = resnet(sd, "encoder.mid_block.resnets.0", h, G)
    h = mid_attention(sd, "encoder.mid_block.attentions.0", h, G)
    h = resnet(sd, "encoder.mid_block.resnets.1", h, G)
    h = F.silu(F.group_norm(h, G, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], eps=1e-6))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def vae_decode(sd, cfg: VaeConfig, z):
    """z (n, latent_channels, h, w) already divided by scaling_factor -> (n, 3, 8h, 8w) for the 4-block SD-VAE."""
    G = cfg.norm_num_groups
    x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = F.conv2d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = resnet(sd, "decoder.mid_block.resnets.0", x, G)
    x = mid_attention(sd, "decoder.mid_block.attentions.0", x, G)
    x = resnet(sd, "decoder.mid_block.resnets.1", x, G)
    up = cfg.up_channels
    for b in range(len(up)):
        for r in range(cfg.layers_per_block + 1):
            x = resnet(sd, f"decoder.up_blocks.{b}.resnets.{r}", x, G)
        if b + 1 < len(up):
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"], sd[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(F.group_norm(x, G, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], eps=1e-6))
    return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


# ==================================================================================================================
# AutoencoderKLTemporalDecoder.decode(z, num_frames) — the SVD temporal decoder the T2V pipeline uses in chunks of 14
# frames (sample/pipeline_latte.py:779-798).  PARITY UNPINNED, restated from diffusers 0.24.0:
# every resnet is a SpatioTemporalResBlock = ResnetBlock2D (eps 1e-6) followed by a TemporalResnetBlock
# (GroupNorm over (c/g, f, h, w), SiLU, Conv3d (3,1,1) pad (1,0,0), twice, + input; eps 1e-5) and blended by
# AlphaBlender(merge_strategy="learned", switch_spatial_to_temporal_mix=True):
#     alpha = 1 - sigmoid(mix_factor);  out = alpha * x_spatial + (1 - alpha) * x_temporal
# then GroupNorm/SiLU/conv_out and time_conv_out = Conv3d(3, 3, (3,1,1), pad (1,0,0)).  No post_quant_conv.
def temporal_state_dict_spec(cfg: VaeConfig):
    up = cfg.up_channels
    C0 = up[0]

    def st(prefix, cin, cout):
        s = _resnet_spec(prefix + ".spatial_res_block", cin, cout)
        t = prefix + ".temporal_res_block"
        s += [(f"{t}.norm1.weight", (cout,)), (f"{t}.norm1.bias", (cout,)), (f"{t}.conv1.weight", (cout, cout, 3, 1, 1)), (f"{t}.conv1.bias", (cout,)),
              (f"{t}.norm2.weight", (cout,)), (f"{t}.norm2.bias", (cout,)), (f"{t}.conv2.weight", (cout, cout, 3, 1, 1)), (f"{t}.conv2.bias", (cout,)),
              (f"{prefix}.time_mixer.mix_factor", (1,))]
        return s

    spec = [("decoder.conv_in.weight", (C0, cfg.latent_channels, 3, 3)), ("decoder.conv_in.bias", (C0,))]
    spec += st("decoder.mid_block.resnets.0", C0, C0)
    a = "decoder.mid_block.attentions.0"
    spec += [(f"{a}.group_norm.weight", (C0,)), (f"{a}.group_norm.bias", (C0,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        spec += [(f"{a}.{n}.weight", (C0, C0)), (f"{a}.{n}.bias", (C0,))]
    spec += st("decoder.mid_block.resnets.1", C0, C0)
    cin = C0
    for b, co in enumerate(up):
        for r in range(cfg.layers_per_block + 1):
            spec += st(f"decoder.up_blocks.{b}.resnets.{r}", cin if r == 0 else co, co)
        if b + 1 < len(up):
            spec += [(f"decoder.up_blocks.{b}.upsamplers.0.conv.weight", (co, co, 3, 3)), (f"decoder.up_blocks.{b}.upsamplers.0.conv.bias", (co,))]
        cin = co
    spec += [("decoder.conv_norm_out.weight", (up[-1],)), ("decoder.conv_norm_out.bias", (up[-1],)),
             ("decoder.conv_out.weight", (cfg.out_channels, up[-1], 3, 3)), ("decoder.conv_out.bias", (cfg.out_channels,)),
             ("decoder.time_conv_out.weight", (cfg.out_channels, cfg.out_channels, 3, 1, 1)), ("decoder.time_conv_out.bias", (cfg.out_channels,))]
    return spec


def make_temporal_weights(cfg: VaeConfig, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in temporal_state_dict_spec(cfg):
        if name.endswith("mix_factor"):
            t = torch.randn(shape, generator=g)
        elif "norm" in name.split(".")[-2] and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = t.contiguous()
    return sd


def spatio_temporal_block(sd, p, x, groups, num_frames):
    xs = resnet(sd, p + ".spatial_res_block", x, groups)
    n, c, h, w = xs.shape
    b = n // num_frames
    v = xs.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)                    # b c f h w
    t = p + ".temporal_res_block"
    u = F.conv3d(F.silu(F.group_norm(v, groups, sd[t + ".norm1.weight"], sd[t + ".norm1.bias"], eps=1e-5)), sd[t + ".conv1.weight"], sd[t + ".conv1.bias"], padding=(1, 0, 0))
    u = F.conv3d(F.silu(F.group_norm(u, groups, sd[t + ".norm2.weight"], sd[t + ".norm2.bias"], eps=1e-5)), sd[t + ".conv2.weight"], sd[t + ".conv2.bias"], padding=(1, 0, 0))
    xt = v + u
    alpha = 1.0 - torch.sigmoid(sd[p + ".time_mixer.mix_factor"])
    out = alpha * v + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(n, c, h, w)


def vae_temporal_decode(sd, cfg: VaeConfig, z, num_frames):
    G = cfg.norm_num_groups
    x = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = spatio_temporal_block(sd, "decoder.mid_block.resnets.0", x, G, num_frames)
    x = mid_attention(sd, "decoder.mid_block.attentions.0", x, G)
    x = spatio_temporal_block(sd, "decoder.mid_block.resnets.1", x, G, num_frames)
    up = cfg.up_channels
    for b in range(len(up)):
        for r in range(cfg.layers_per_block + 1):
            x = spatio_temporal_block(sd, f"decoder.up_blocks.{b}.resnets.{r}", x, G, num_frames)
        if b + 1 < len(up):
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"], sd[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(F.group_norm(x, G, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], eps=1e-6))
    x = F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
    n, c, h, w = x.shape
    v = x.reshape(n // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    v = F.conv3d(v, sd["decoder.time_conv_out.weight"], sd["decoder.time_conv_out.bias"], padding=(1, 0, 0))
    return v.permute(0, 2, 1, 3, 4).reshape(n, c, h, w)
