"""GPU: LatteT2V forward through the C ABI against the CPU oracle restatement (oracle/t2v_oracle.py).
PARITY UNPINNED: the oracle restates diffusers 0.24.0 pieces that are not available offline (see its header), so these
tests show agreement with the restatement only.  Tolerance: 1e-2 max-abs on O(5) outputs with fp16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cfg_kw, batch, text_len, seed, temporal=True):
    from latte_b200 import LatteT2V
    from oracle import t2v_oracle as T
    cfg = T.T2VConfig(**cfg_kw)
    sd = T.make_weights(cfg, seed)
    x, t, text = T.make_inputs(cfg, batch, text_len, seed + 1)
    net = LatteT2V(**cfg_kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda(), t.cuda(), encoder_hidden_states=text.cuda(), enable_temporal_attentions=temporal, return_dict=False)[0]
    ref = T.t2v_forward(sd, cfg, x, t, text, enable_temporal=temporal)
    assert out.shape == ref.shape == (batch, cfg.out_channels, cfg.video_length, cfg.sample_size, cfg.sample_size)
    return (out.cpu() - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("case", [
    (dict(num_attention_heads=2, attention_head_dim=64, num_layers=2, sample_size=16, video_length=8, caption_channels=256), 2, 20),
    (dict(num_attention_heads=8, attention_head_dim=72, num_layers=2, sample_size=32, video_length=16, caption_channels=512), 2, 120),
    (dict(num_attention_heads=8, attention_head_dim=72, num_layers=1, sample_size=64, video_length=4, caption_channels=256), 1, 33),
])
def test_t2v_forward_matches_oracle(case):
    kw, batch, L = case
    err, mag = _run(kw, batch, L, 3)
    assert err < 1e-2, f"max-abs {err:.3e} (output magnitude {mag:.2f})"


def test_t2v_without_temporal_blocks():
    """enable_temporal_attentions=False (T2I path, pipeline_latte.py:706): only the spatial blocks run."""
    kw = dict(num_attention_heads=2, attention_head_dim=64, num_layers=2, sample_size=16, video_length=8, caption_channels=256)
    err, _ = _run(kw, 2, 16, 5, temporal=False)
    assert err < 1e-2


def test_t2v_surface():
    from latte_b200 import LatteT2V
    net = LatteT2V(num_attention_heads=2, attention_head_dim=64, num_layers=1, sample_size=16, video_length=8, caption_channels=256)
    assert net.config.in_channels == 4 and net.config.out_channels == 8 and net.config.sample_size == 16
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 4, 8, 16, 16), torch.tensor([1]), encoder_hidden_states=torch.zeros(1, 8, 256))
    net = net.cuda().half()
    assert net.dtype == torch.float16
    out = net(torch.randn(1, 4, 8, 16, 16, device="cuda").half(), torch.tensor([5], device="cuda"),
              encoder_hidden_states=torch.randn(1, 8, 256, device="cuda").half())
    assert out.sample.dtype == torch.float16 and out.sample.shape == (1, 8, 8, 16, 16)
