"""diffusers.models.attention.BasicTransformerBlock (shim restatement of 0.24.0), the SPATIAL block of LatteT2V
(constructed at latte_t2v.py:587-603, called at :862-870), for norm_type 'ada_norm_single' only:

    six chunks of scale_shift_table[None] + timestep.reshape(B, 6, D)
    h += gate_msa * attn1(norm1(h) * (1 + scale_msa) + shift_msa)
    h += attn2(h, encoder_hidden_states, encoder_attention_mask)        (no norm before attn2 in this mode)
    h += gate_mlp * ff(norm2(h) * (1 + scale_mlp) + shift_mlp)

The reference's own temporal variant `BasicTransformerBlock_` (latte_t2v.py:126-396) is an edited copy of this class and
is NOT restated here -- it runs from the reference file."""
import torch
import torch.nn as nn

from .activations import GELU
from .attention_processor import Attention
from .lora import LoRACompatibleLinear


class FeedForward(nn.Module):
    def __init__(self, dim, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        if activation_fn not in ("gelu", "gelu-approximate"):
            raise NotImplementedError("diffusers shim: only the gelu feed-forwards are restated")
        act = GELU(dim, 4 * dim, approximate="tanh" if activation_fn == "gelu-approximate" else "none")
        self.net = nn.ModuleList([act, nn.Dropout(dropout), LoRACompatibleLinear(4 * dim, dim)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, upcast_attention=False, norm_elementwise_affine=True,
                 norm_type="layer_norm", norm_eps=1e-5, final_dropout=False, attention_type="default", **unused):
        super().__init__()
        if norm_type != "ada_norm_single" or only_cross_attention or double_self_attention or attention_type != "default":
            raise NotImplementedError("diffusers shim: BasicTransformerBlock is restated for ada_norm_single only")
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine, eps=norm_eps)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias, upcast_attention=upcast_attention)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine, eps=norm_eps)
            self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                                   dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                   upcast_attention=upcast_attention)
        else:
            self.norm2 = None
            self.attn2 = None
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        batch = hidden_states.shape[0]
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (
            self.scale_shift_table[None] + timestep.reshape(batch, 6, -1)).chunk(6, dim=1)
        h = self.norm1(hidden_states) * (1 + scale_msa) + shift_msa
        hidden_states = gate_msa * self.attn1(h, attention_mask=attention_mask) + hidden_states
        if self.attn2 is not None:
            hidden_states = self.attn2(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                       attention_mask=encoder_attention_mask) + hidden_states
        h = self.norm2(hidden_states) * (1 + scale_mlp) + shift_mlp
        return gate_mlp * self.ff(h) + hidden_states
