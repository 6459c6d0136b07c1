"""GPU: AutoencoderKL.decode through the C ABI (TMA implicit-GEMM convolutions) against the CPU oracle restatement.
PARITY UNPINNED (diffusers 0.24.0 absent offline; see oracle/vae_oracle.py).  Activations are 16-bit NHWC through ~30
layers like the reference's `vae.half()` path: tolerance 3e-2 max-abs / 3e-3 mean-abs on O(1) outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(block_out, groups, n, h, w, seed):
    from latte_b200 import AutoencoderKL
    from oracle import vae_oracle as V
    cfg = V.VaeConfig(block_out_channels=block_out, norm_num_groups=groups)
    sd = V.make_weights(cfg, seed)
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn(n, 4, h, w, generator=g)
    vae = AutoencoderKL(block_out_channels=block_out, norm_num_groups=groups)
    vae.load_state_dict(sd, strict=True)
    vae = vae.cuda().eval()
    with torch.no_grad():
        out = vae.decode(z.cuda()).sample.cpu()
        ref = V.vae_decode(sd, cfg, z)
    scale = 2 ** (len(block_out) - 1)
    assert out.shape == ref.shape == (n, 3, h * scale, w * scale)
    err = (out - ref).abs()
    return err.max().item(), err.mean().item(), ref.abs().max().item()


@pytest.mark.parametrize("case", [
    ((64, 128, 128), 16, 2, 16, 16),        # 3 blocks: 16x16 -> 64x64, packed-row tiles (W < 128)
    ((64, 64, 128, 128), 16, 3, 16, 16),    # 4 blocks -> 128x128: exercises W = 128 tiles
    ((128, 256, 512, 512), 32, 1, 32, 32),  # the SD-VAE topology at the BASELINE latent size (one frame)
])
def test_decode_matches_oracle(case):
    mx, mean, mag = _run(*case, seed=3)
    assert mx < 3e-2 and mean < 3e-3, f"max {mx:.3e} mean {mean:.3e} (|ref| max {mag:.2f})"


def test_surface():
    from latte_b200 import AutoencoderKL
    vae = AutoencoderKL(block_out_channels=(64, 128, 128), norm_num_groups=16)
    assert vae.config.scaling_factor == 0.18215 and vae.config.block_out_channels == (64, 128, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vae.decode(torch.zeros(1, 4, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vae.encode(torch.zeros(1, 3, 64, 64))
    vae = vae.cuda().half()
    out = vae.decode(torch.randn(2, 4, 16, 16, device="cuda").half(), num_frames=2).sample   # kwargs like the temporal decoder call
    assert out.dtype == torch.float16 and out.shape == (2, 3, 64, 64)


@pytest.mark.parametrize("case", [((64, 128, 128), 16, 4, 16, 16), ((64, 64, 128, 128), 16, 14, 16, 16)])
def test_temporal_decoder_matches_oracle(case):
    """AutoencoderKLTemporalDecoder.decode(z, num_frames) (pipeline_latte.py:779-798; 14-frame chunk): spatial resnets +
    Conv3d(3,1,1) temporal resnets (GroupNorm over the clip) blended by the learned alpha, then time_conv_out."""
    from latte_b200 import AutoencoderKLTemporalDecoder
    from oracle import vae_oracle as V
    block_out, groups, n, h, w = case
    cfg = V.VaeConfig(block_out_channels=block_out, norm_num_groups=groups)
    sd = V.make_temporal_weights(cfg, 5)
    z = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(6))
    vae = AutoencoderKLTemporalDecoder(block_out_channels=block_out, norm_num_groups=groups)
    vae.load_state_dict(sd, strict=True)
    vae = vae.cuda().eval()
    with torch.no_grad():
        out = vae.decode(z.cuda(), num_frames=n).sample.cpu()
        ref = V.vae_temporal_decode(sd, cfg, z, n)
    err = (out - ref).abs()
    assert out.shape == ref.shape
    assert err.max().item() < 3e-2 and err.mean().item() < 3e-3, f"max {err.max().item():.3e} mean {err.mean().item():.3e}"
    with pytest.raises(RuntimeError, match="ONE clip"):
        vae.decode(z.cuda(), num_frames=n // 2)


@pytest.mark.parametrize("case", [
    ((64, 128, 128), 16, 2, 128, 128),       # 3 blocks: 128x128 -> 32x32 latent (two stride-2 convs, W = 128 and packed-row tiles)
    ((64, 64, 128, 128), 16, 3, 128, 128),   # 4 blocks -> 16x16 latent
    ((128, 256, 512, 512), 32, 1, 256, 256), # the SD-VAE topology at train.py's 256x256 frames
])
def test_encode_matches_oracle(case):
    """AutoencoderKL.encode(x).latent_dist (train.py:206-211): moments = quant_conv(Encoder(x)) against the CPU oracle
    restatement (PARITY UNPINNED, like decode); the stride-2 convolutions run as space-to-depth + 2x2-tap implicit GEMMs."""
    from latte_b200 import AutoencoderKL
    from oracle import vae_oracle as V
    block_out, groups, n, h, w = case
    cfg = V.VaeConfig(block_out_channels=block_out, norm_num_groups=groups)
    sd = V.make_weights(cfg, 9)
    g = torch.Generator().manual_seed(10)
    x = torch.rand(n, 3, h, w, generator=g) * 2 - 1
    vae = AutoencoderKL(block_out_channels=block_out, norm_num_groups=groups)
    vae.load_state_dict(sd, strict=True)
    vae = vae.cuda().eval()
    with torch.no_grad():
        dist = vae.encode(x.cuda()).latent_dist
        ref = V.vae_encode(sd, cfg, x)
    f = 2 ** (len(block_out) - 1)
    got = dist.parameters.cpu()
    assert got.shape == ref.shape == (n, 8, h // f, w // f)
    err = (got - ref).abs()
    assert err.max().item() < 3e-2 and err.mean().item() < 3e-3, f"max {err.max():.3e} mean {err.mean():.3e} (|ref| max {ref.abs().max():.2f})"
    # DiagonalGaussianDistribution surface: mean | clamped logvar, reparameterised sample with the caller's generator
    assert torch.equal(dist.mean, dist.parameters[:, :4]) and torch.equal(dist.mode(), dist.mean)
    gen = torch.Generator(device="cuda").manual_seed(1)
    s1 = dist.sample(generator=gen)
    gen.manual_seed(1)
    noise = torch.randn(dist.mean.shape, generator=gen, device="cuda")
    assert torch.allclose(s1, dist.mean + dist.std * noise)
    assert dist.kl().shape == (n,)


def test_encode_then_decode_shapes():
    from latte_b200 import AutoencoderKL
    vae = AutoencoderKL(block_out_channels=(64, 128, 128), norm_num_groups=16).cuda().half()
    x = torch.rand(2, 3, 128, 128, device="cuda").half() * 2 - 1
    z = vae.encode(x).latent_dist.sample().mul_(0.18215)             # train.py:210
    assert z.dtype == torch.float16 and z.shape == (2, 4, 32, 32)
    out = vae.decode(z / 0.18215).sample
    assert out.shape == (2, 3, 128, 128) and torch.isfinite(out).all()
