"""diffusers.models.attention_processor.Attention with the default AttnProcessor2_0 (shim restatement of 0.24.0).

to_q / to_k / to_v / to_out[0] Linear (+ Dropout at to_out[1]); heads split; torch scaled_dot_product_attention with
scale head_dim**-0.5 and an optional ADDITIVE mask; heads merged; output projection.  A mask arrives as a bias of shape
(batch, 1, key_tokens) (latte_t2v.py:752-771); `prepare_attention_mask` repeats it per head and the processor views it
as (batch, heads, 1, key_tokens) so that it broadcasts over the query tokens."""
import torch.nn as nn
import torch.nn.functional as F

from .lora import LoRACompatibleLinear


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, out_bias=True, **unused):
        super().__init__()
        inner = dim_head * heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = LoRACompatibleLinear(query_dim, inner, bias=bias)
        self.to_k = LoRACompatibleLinear(kv_dim, inner, bias=bias)
        self.to_v = LoRACompatibleLinear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(inner, query_dim, bias=out_bias), nn.Dropout(dropout)])

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, scale: float = 1.0, **unused):
        context = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        batch, key_tokens, _ = context.shape
        if attention_mask is not None:
            attention_mask = self.prepare_attention_mask(attention_mask, key_tokens, batch)
            attention_mask = attention_mask.view(batch, self.heads, -1, attention_mask.shape[-1])
        q = self.to_q(hidden_states)
        k = self.to_k(context)
        v = self.to_v(context)
        head_dim = k.shape[-1] // self.heads
        q = q.view(batch, -1, self.heads, head_dim).transpose(1, 2)
        k = k.view(batch, -1, self.heads, head_dim).transpose(1, 2)
        v = v.view(batch, -1, self.heads, head_dim).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, self.heads * head_dim).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))
