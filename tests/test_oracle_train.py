"""CPU: the oracle's BACKWARD is pinned too.  One training step (diffusion.training_losses on Latte, train.py:206-222) run
through the oracle restatements with autograd must reproduce the loss and the parameter gradients that the unmodified
reference produced (tests/golden/train_tiny64.npz, oracle/make_golden_train.py).  Groundwork for the training row
(BASELINE config 5): no backward kernel exists yet, the product raises under grad (tests/test_module.py)."""
import os

import numpy as np
import torch

from oracle import latte_oracle as O
from oracle import sampler_oracle as S


def test_training_step_gradients_equal_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_tiny64.npz"))
    cfg = O.make_config("Latte-tiny64/2", input_size=16, num_frames=8)
    sd = {k: v.clone().requires_grad_(k not in ("pos_embed", "temp_embed")) for k, v in O.make_weights(cfg, 21).items()}
    x0, noise = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"])
    t, y = torch.from_numpy(g["t"]), torch.from_numpy(g["y"])
    s = S.make_schedule("")
    terms = S.training_losses(s, lambda x, tt, **kw: O.latte_forward(sd, cfg, x, tt, kw["y"]), x0, t, noise, dict(y=y))
    loss = terms["loss"].mean()
    # forward: same fp32 ops on the same inputs; the functional restatement may associate a few sums differently
    np.testing.assert_allclose(np.stack([terms[k].detach().numpy() for k in ("loss", "mse", "vb")]), g["loss_terms"], rtol=2e-5, atol=1e-6)
    assert abs(loss.item() - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    loss.backward()
    names, norms = list(g["grad_names"]), g["grad_norms"]
    assert len(names) == sum(1 for v in sd.values() if v.requires_grad and v.grad is not None)
    for k, want in zip(names, norms):
        got = sd[k].grad.double().norm().item()
        assert abs(got - want) <= 1e-4 * want + 1e-9, (k, got, want)
    for key in g.files:
        if key.startswith("grad::"):
            k = key[len("grad::"):]
            ref = torch.from_numpy(g[key])
            err = (sd[k].grad - ref).abs().max().item()
            assert err <= 1e-4 * ref.abs().max().item() + 1e-8, (k, err)
