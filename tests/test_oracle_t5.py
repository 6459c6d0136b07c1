"""CPU: the T5 encoder restatement (oracle/t5_oracle.py) against fixtures produced by transformers' own T5EncoderModel
(oracle/make_golden_t5.py) -- the library the reference's pipeline calls for prompt embeddings (pipeline_latte.py:214)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import t5_oracle as T


def load_case(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"t5_{tag}.npz"))
    cfg = T.T5Cfg(**ast.literal_eval(str(g["cfg"])))
    sd = T.make_weights(cfg, int(g["wseed"]))
    return g, cfg, sd, torch.from_numpy(g["ids"]), torch.from_numpy(g["mask"])


@pytest.mark.parametrize("tag", ["tiny_b2_l20", "wide_b2_l120", "wide_b1_l128"])
def test_oracle_matches_transformers_golden(golden_dir, tag):
    g, cfg, sd, ids, mask = load_case(golden_dir, tag)
    out = T.t5_encode(sd, cfg, ids, mask)
    ref = torch.from_numpy(g["out"])
    keep = mask.bool()
    # rows of masked (padding) tokens are still computed by the library; they attend to the kept tokens only, like here
    assert (out - ref).abs().max().item() < 2e-4
    assert keep.any()


def test_bucket_function_matches_the_module():
    from latte_b200.t5 import relative_position_buckets
    pos = torch.arange(128)
    assert torch.equal(relative_position_buckets(128), T.relative_position_bucket(pos[None, :] - pos[:, None], 32, 128))
