"""GPU: the C-ABI forward is CUDA-graph capturable -- no allocation, no synchronisation, no host-side
launch counter inside the call -- and a replay is bit-identical to the eager call, including across the ordered
stream-K GEMMs whose flags must be left clean by every launch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(name, **kw):
    from latte_b200 import Latte
    from oracle import latte_oracle as O
    cfg = O.make_config(name, **kw)
    sd = O.make_weights(cfg, 5)
    x, t, y = O.make_inputs(cfg, 2, 6)
    net = Latte(input_size=cfg.input_size, hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads,
                num_frames=cfg.num_frames, num_classes=cfg.num_classes, learn_sigma=True, extras=2)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), x.cuda(), t.cuda(), y.cuda()


@pytest.mark.parametrize("name,kw", [("Latte-tiny72/2", dict(input_size=16, num_frames=16)), ("Latte-XL/2", {})])
def test_forward_graph_replay_is_bit_identical(name, kw):
    net, x, t, y = _net(name, **kw)
    net.use_cuda_graphs = False          # capture the raw C-ABI call ourselves
    with torch.no_grad():
        eager = net.forward_with_cfg(x, t, y=y, cfg_scale=7.0).clone()      # also warms attributes / descriptor cache
        xs = x.clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net.forward_with_cfg(xs, t, y=y, cfg_scale=7.0)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = net.forward_with_cfg(xs, t, y=y, cfg_scale=7.0)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
        xs.copy_(x * 0.5)                 # new input through the same graph
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, net.forward_with_cfg(x * 0.5, t, y=y, cfg_scale=7.0))


def test_module_graph_mode_matches_eager():
    """`Latte.use_cuda_graphs` (default on in eval / no_grad): the module captures its own forward per (batch, cfg) key
    and replays it; results are bit-identical to the eager path and inputs may change between calls."""
    net, x, t, y = _net("Latte-tiny72/2", input_size=16, num_frames=16)
    with torch.no_grad():
        net.use_cuda_graphs = False
        ref1 = net.forward_with_cfg(x, t, y=y, cfg_scale=7.0).clone()
        ref2 = net(x * 0.25, t + 3, y=y).clone()
        net.use_cuda_graphs = True
        for _ in range(3):                # 1st call eager (warm-up), 2nd captures, 3rd replays
            got1 = net.forward_with_cfg(x, t, y=y, cfg_scale=7.0).clone()
        for _ in range(3):
            got2 = net(x * 0.25, t + 3, y=y).clone()
        got1b = net.forward_with_cfg(x, t, y=y, cfg_scale=7.0)
    assert torch.equal(got1, ref1) and torch.equal(got2, ref2) and torch.equal(got1b, ref1)
