"""CPU: the ORCHESTRATION of the training step (latte_b200/training.py: which activations are kept, the chain rule over the
blocks, the per-sample adaLN reductions, the autograd boundary that hands every parameter its gradient) is exact.  The engine is
driven through oracle/train_ops_oracle.TorchOps in fp32 -- the torch restatement of the ops the CUDA backend implements -- under
the product's own `diffusion.training_losses`, and must reproduce the loss and the parameter gradients that the UNMODIFIED
reference produced (tests/golden/train_tiny64.npz, oracle/make_golden_train.py: models/latte.py under
diffusion.training_losses, train.py:206-222).  The kernels themselves are compared with the same TorchOps on the GPU
(tests/test_gpu_train.py)."""
import os

import numpy as np
import torch

from latte_b200 import Latte, training
from latte_b200.diffusion import create_diffusion
from oracle import latte_oracle as O
from oracle.train_ops_oracle import TorchOps


def _setup(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_tiny64.npz"))
    cfg = O.make_config("Latte-tiny64/2", input_size=16, num_frames=8)
    sd = O.make_weights(cfg, 21)
    m = Latte(input_size=16, hidden_size=128, depth=2, num_heads=2, num_frames=8, num_classes=101, extras=2)
    m.load_state_dict(sd, strict=True)
    return g, cfg, sd, m


def test_training_losses_match_reference_terms(golden_dir):
    """latte_b200.diffusion.training_losses (q_sample, MSE, learned-range VB with the t == 0 decoder NLL) on the oracle forward."""
    g, cfg, sd, _ = _setup(golden_dir)
    d = create_diffusion(timestep_respacing="")
    x0, noise = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"])
    t, y = torch.from_numpy(g["t"]), torch.from_numpy(g["y"])
    with torch.no_grad():
        terms = d.training_losses(lambda x, tt, **kw: O.latte_forward(sd, cfg, x, tt, kw["y"]), x0, t, dict(y=y), noise=noise)
    got = np.stack([terms[k].numpy() for k in ("loss", "mse", "vb")])
    np.testing.assert_allclose(got, g["loss_terms"], rtol=2e-5, atol=1e-6)


def test_engine_gradients_equal_reference(golden_dir):
    g, cfg, sd, m = _setup(golden_dir)
    m.eval()                     # the golden was made in eval mode (no label dropout RNG); the engine itself has no mode
    d = create_diffusion(timestep_respacing="")
    x0, noise = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"])
    t, y = torch.from_numpy(g["t"]), torch.from_numpy(g["y"])
    ops = TorchOps(torch.float32)

    def model_fn(x, tt, y):
        return training.train_forward(m, ops, torch.float32, x, training.conditioning(m, tt, y))

    terms = d.training_losses(model_fn, x0, t, dict(y=y), noise=noise)
    loss = terms["loss"].mean()          # train.py:222
    assert abs(loss.item() - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    loss.backward()
    named = dict(m.named_parameters())
    names = [str(k) for k in g["grad_names"]]
    assert set(names) == {k for k, p in named.items() if p.grad is not None}
    for k, want in zip(names, g["grad_norms"]):
        got = named[k].grad.double().norm().item()
        assert abs(got - want) <= 1e-4 * want + 1e-9, (k, got, want)
    for key in g.files:
        if key.startswith("grad::"):
            ref = torch.from_numpy(g[key])
            err = (named[key[6:]].grad - ref).abs().max().item()
            assert err <= 1e-4 * ref.abs().max().item() + 1e-8, (key, err)


def test_backward_is_single_use(golden_dir):
    g, cfg, sd, m = _setup(golden_dir)
    x = torch.from_numpy(g["x0"])
    out = training.train_forward(m, TorchOps(torch.float32), torch.float32, x,
                                 training.conditioning(m, torch.from_numpy(g["t"]), torch.from_numpy(g["y"])))
    out.sum().backward(retain_graph=True)
    try:
        out.sum().backward()
    except RuntimeError as e:
        assert "backward called twice" in str(e)
    else:
        raise AssertionError("second backward over freed activations must raise")


def test_engine_with_16bit_operands_on_cpu(golden_dir):
    """The same orchestration with bf16 operand rounding at every point where the CUDA backend rounds (operand casts, GEMM /
    attention / LayerNorm outputs): gradients stay within bf16 noise of the reference's fp32 ones -- i.e. no reduction or
    accumulation in the engine runs in 16 bits by accident."""
    g, cfg, sd, m = _setup(golden_dir)
    m.eval()
    d = create_diffusion(timestep_respacing="")
    x0, noise = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"])
    t, y = torch.from_numpy(g["t"]), torch.from_numpy(g["y"])
    ops = TorchOps(torch.bfloat16)
    terms = d.training_losses(lambda x, tt, y: training.train_forward(m, ops, torch.bfloat16, x, training.conditioning(m, tt, y)),
                              x0, t, dict(y=y), noise=noise)
    loss = terms["loss"].mean()
    assert abs(loss.item() - float(g["loss"])) < 2e-2 * abs(float(g["loss"]))
    loss.backward()
    named = dict(m.named_parameters())
    for k, want in zip([str(n) for n in g["grad_names"]], g["grad_norms"]):
        assert named[k].grad.dtype == torch.float32
        got = named[k].grad.double().norm().item()
        assert abs(got - want) <= 8e-2 * want, (k, got, want)
