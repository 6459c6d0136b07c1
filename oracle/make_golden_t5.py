"""Generate tests/golden/t5_*.npz by running transformers' own T5EncoderModel (the reference's text encoder,
sample/pipeline_latte.py:214) on seeded weights / token ids.  TEST INFRASTRUCTURE.  Usage: python oracle/make_golden_t5.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import t5_oracle as T  # noqa: E402


def gen(tag, cfg_kw, batch, length, valid, wseed, iseed, out_dir):
    from transformers import T5Config, T5EncoderModel
    cfg = T.T5Cfg(**cfg_kw)
    sd = T.make_weights(cfg, wseed)
    ids, mask = T.make_inputs(cfg, batch, length, valid, iseed)
    hf = T5EncoderModel(T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff,
                                 num_layers=cfg.num_layers, num_heads=cfg.num_heads, feed_forward_proj="gated-gelu",
                                 relative_attention_num_buckets=cfg.relative_attention_num_buckets,
                                 relative_attention_max_distance=cfg.relative_attention_max_distance,
                                 layer_norm_epsilon=cfg.layer_norm_epsilon, dropout_rate=0.0))
    missing, unexpected = hf.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    with torch.no_grad():
        out = hf.eval()(ids, attention_mask=mask)[0]
    path = os.path.join(out_dir, f"t5_{tag}.npz")
    np.savez_compressed(path, out=out.numpy(), ids=ids.numpy(), mask=mask.numpy(), cfg=np.array(repr(cfg_kw)), wseed=np.int64(wseed))
    print(f"{path}: out {tuple(out.shape)} absmax {out.abs().max():.3f} std {out.std():.3f}")


TINY = dict(vocab_size=100, d_model=256, d_ff=512, num_layers=2, num_heads=4)
WIDE = dict(vocab_size=1000, d_model=512, d_ff=1024, num_layers=3, num_heads=8)

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden")
    gen("tiny_b2_l20", TINY, 2, 20, [7, 20], 11, 12, out_dir)
    gen("wide_b2_l120", WIDE, 2, 120, [120, 33], 13, 14, out_dir)
    gen("wide_b1_l128", WIDE, 1, 128, [128], 13, 15, out_dir)
