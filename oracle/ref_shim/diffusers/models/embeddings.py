"""diffusers.models.embeddings (shim restatement of 0.24.0): the sin-cos tables, PatchEmbed, the PixArt timestep
embedding stack and the caption projection used by LatteT2V (latte_t2v.py:11,17,412-414,575-582,664,943-944)."""
import math

import numpy as np
import torch
import torch.nn as nn


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """[sin | cos] of pos x omega, omega_k = 10000^(-k / (embed_dim/2)), evaluated in float64 numpy."""
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    angles = np.einsum("m,d->md", np.asarray(pos).reshape(-1), omega)
    return np.concatenate([np.sin(angles), np.cos(angles)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, base_size=16, interpolation_scale=1.0):
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])   # w first, as published
    first = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    second = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return np.concatenate([first, second], axis=1)


class PatchEmbed(nn.Module):
    """Conv2d(k = s = patch) -> tokens, plus a fixed 2-D sin-cos table (non-persistent buffer) cast to the input dtype."""

    def __init__(self, height=224, width=224, patch_size=16, in_channels=3, embed_dim=768, bias=True, interpolation_scale=1):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.patch_size = patch_size
        self.height, self.width = height // patch_size, width // patch_size
        self.base_size = height // patch_size
        self.interpolation_scale = interpolation_scale
        num_patches = self.height * self.width
        table = get_2d_sincos_pos_embed(embed_dim, int(num_patches ** 0.5), base_size=self.base_size,
                                        interpolation_scale=self.interpolation_scale)
        self.register_buffer("pos_embed", torch.from_numpy(table).float().unsqueeze(0), persistent=False)

    def forward(self, latent):
        assert latent.shape[-2] // self.patch_size == self.height and latent.shape[-1] // self.patch_size == self.width, \
            "diffusers shim: PatchEmbed at a resolution other than the configured one is not restated"
        latent = self.proj(latent).flatten(2).transpose(1, 2)
        return (latent + self.pos_embed).to(latent.dtype)


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    freqs = torch.exp(exponent / (half - downscale_freq_shift))
    args = timesteps[:, None].float() * freqs[None, :]
    emb = torch.cat([torch.sin(args), torch.cos(args)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class CombinedTimestepSizeEmbeddings(nn.Module):
    """PixArt-alpha conditioning; Latte-1 (sample_size 64) has use_additional_conditions = False, so this is
    Timesteps(256, flip_sin_to_cos=True, shift 0) -> TimestepEmbedding."""

    def __init__(self, embedding_dim, size_emb_dim, use_additional_conditions: bool = False):
        super().__init__()
        if use_additional_conditions:
            raise NotImplementedError("diffusers shim: resolution / aspect-ratio conditioning is not restated")
        self.time_proj = Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)
        self.timestep_embedder = TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)

    def forward(self, timestep, resolution=None, aspect_ratio=None, batch_size=None, hidden_dtype=None):
        return self.timestep_embedder(self.time_proj(timestep).to(dtype=hidden_dtype))


class CaptionProjection(nn.Module):
    def __init__(self, in_features, hidden_size, num_tokens=120):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.act_1 = nn.GELU(approximate="tanh")
        self.linear_2 = nn.Linear(hidden_size, hidden_size)
        self.register_buffer("y_embedding", nn.Parameter(torch.randn(num_tokens, in_features) / in_features ** 0.5))

    def forward(self, caption, force_drop_ids=None):
        assert force_drop_ids is None
        return self.linear_2(self.act_1(self.linear_1(caption)))


class _Unbuilt(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"diffusers shim: {type(self).__name__} is a placeholder (not used by Latte-1)")


class ImagePositionalEmbeddings(_Unbuilt):
    pass


class SinusoidalPositionalEmbedding(_Unbuilt):
    pass
