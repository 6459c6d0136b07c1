"""Training step of `Latte` (BASELINE config 5: train.py:206-222 — forward + backward under `loss.backward()`).

The reference trains through torch autograd over ~1000 eager kernels per forward.  Here the forward keeps the activations the
backward needs and the backward is written out explicitly, op by op, over the same hand-written kernels as the sampling path
(the tcgen05 GEMM does every dgrad and wgrad; wgrads accumulate in fp32 straight into the gradient buffers through the
residual epilogue) plus the kernels of csrc/train.cu (LayerNorm-modulate backward, gate / GELU backward with the bias and
per-sample reductions fused, attention backward on tensor cores, 16-bit transposes, adaLN outer products).

Derivatives follow the reference forward (models/latte.py): block :177-181, modulate :28-29, attention 'math' :48-77, Mlp
:169-171, FinalLayer :197-201, PatchEmbed + pos_embed :330-331, temp_embed :357-358.  Rows stay in (b, f, n) order for
spatial AND temporal blocks (the regrouping of :355/:368 is index arithmetic inside the attention kernels), so every
per-sample adaLN vector addresses `rows_per_batch = F*N` consecutive rows.

`TrainEngine` is backend-agnostic: the product backend is latte_b200.train_ops.NativeOps (C ABI, CUDA only, raises without
the extension); tests drive the same orchestration through oracle/train_ops_oracle.TorchOps on the CPU and compare with
gradients produced by the unmodified reference (tests/golden/train_tiny64.npz).
"""
from __future__ import annotations

import os

import torch

_BLOCK_LINEARS = ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")
_FUSED_RESIDUAL_LN = os.environ.get("LATTE_B200_FUSED_RESIDUAL_LN", "0") == "1"


class TrainEngine:
    def __init__(self, model, ops, dtype):
        self.m = model
        self.ops = ops
        self.dtype = dtype
        self.saved = None
        self.w = None

    # ---------------------------------------------------------------------------------------------------------------
    def prepare(self):
        """Operand copies of the current parameters in the compute type -- ONE copy per weight: the forward GEMM reads it as
        [N][K], dgrad reads the same memory as an MN-major operand.  The 16-bit buffers persist on the model between steps
        (`model._train_operands`); a step refreshes all of them with one multi-tensor cast launch.  adaLN weights of all blocks
        + final layer land in one stacked buffer; patch-embed / final-layer operands are zero-padded to the GEMM's 64-element
        k-block (K = 16 and 32)."""
        m, ops = self.m, self.ops
        D = m.hidden_size
        dev = m.pos_embed.device
        lin = lambda blk, name: (getattr(getattr(blk, name.split(".")[0]), name.split(".")[1]))   # noqa: E731
        ada = [b.adaLN_modulation[1] for b in m.blocks] + [m.final_layer.adaLN_modulation[1]]
        cache = getattr(m, "_train_operands", None)
        key = (self.dtype, dev, type(ops).__name__)
        if cache is None or cache["key"] != key:
            NA = sum(a.weight.shape[0] for a in ada)
            cache = {"key": key, "ada_w": torch.empty(NA, D, dtype=self.dtype, device=dev), "w": {}}
            for i, blk in enumerate(m.blocks):
                for name in _BLOCK_LINEARS:
                    cache["w"][f"{i}.{name}"] = torch.empty(lin(blk, name).weight.shape, dtype=self.dtype, device=dev)
            m._train_operands = cache
        srcs, dsts = [], []
        for i, blk in enumerate(m.blocks):
            for name in _BLOCK_LINEARS:
                srcs.append(lin(blk, name).weight.detach())
                dsts.append(cache["w"][f"{i}.{name}"])
        row = 0
        for a in ada:
            srcs.append(a.weight.detach())
            dsts.append(cache["ada_w"][row:row + a.weight.shape[0]])
            row += a.weight.shape[0]
        if all(t.dtype == torch.float32 and t.is_contiguous() for t in srcs):
            ops.cast_into(srcs, dsts)
        else:                                   # 16-bit or non-contiguous parameters: plain copies
            for a, b in zip(srcs, dsts):
                b.copy_(a)
        W = {}
        for i, blk in enumerate(m.blocks):
            for name in _BLOCK_LINEARS:
                W[f"{i}.{name}"] = (cache["w"][f"{i}.{name}"], lin(blk, name).bias.detach().float().contiguous())
        W["ada_w"] = cache["ada_w"]
        W["ada_b"] = torch.cat([a.bias.detach() for a in ada]).float().contiguous()
        pw = m.x_embedder.proj.weight.detach().reshape(D, -1).float()
        self.kp = pw.shape[1]
        pad = torch.zeros(D, 64, dtype=torch.float32, device=dev)
        pad[:, : self.kp] = pw
        W["patch_w"] = ops.cast(pad)
        W["patch_b"] = m.x_embedder.proj.bias.detach().float().contiguous()
        fw = m.final_layer.linear.weight.detach().float()             # [p*p*Cout, D]
        self.nf = fw.shape[0]
        padk = torch.zeros(64, D, dtype=torch.float32, device=dev)
        padk[: self.nf] = fw
        W["final_wk"] = ops.cast(padk)                                # rows [0, nf) = the weight (forward), all 64 rows = dgrad operand
        W["final_w"] = W["final_wk"][: self.nf]
        W["final_b"] = m.final_layer.linear.bias.detach().float().contiguous()
        self.w = W

    # ---------------------------------------------------------------------------------------------------------------
    def _patchify(self, x):
        """(B, F, C, H, W) -> rows (b, f, gh, gw) x columns (c, i, j): timm PatchEmbed's Conv2d(k = s = p) as a GEMM operand."""
        m = self.m
        B, Fr, C, H, Wd = x.shape
        p = m.patch_size
        xx = x.reshape(B * Fr, C, H // p, p, Wd // p, p).permute(0, 2, 4, 1, 3, 5)
        return xx.reshape(B * Fr * (H // p) * (Wd // p), C * p * p)

    def _unpatchify(self, tok, B):
        """rows (b, f, h, w) x (p, q, c) -> (B, F, c, h*p, w*q) (latte.py:297-310, :375-376)."""
        m = self.m
        c, p = m.out_channels, m.patch_size
        g = m.input_size // p
        t = tok.view(B * m.num_frames, g, g, p, p, c).permute(0, 5, 1, 3, 2, 4)
        return t.reshape(B, m.num_frames, c, g * p, g * p)

    def _patchify_out(self, dout):
        m = self.m
        c, p = m.out_channels, m.patch_size
        g = m.input_size // p
        B = dout.shape[0]
        t = dout.reshape(B * m.num_frames, c, g, p, g, p).permute(0, 2, 4, 3, 5, 1)
        return t.reshape(B * m.num_frames * g * g, p * p * c).contiguous()

    # ---------------------------------------------------------------------------------------------------------------
    def forward(self, x, c):
        """x (B, F, C, H, W) fp32, c (B, D) fp32 = t_embedder(t) + y_embedder(y) (latte.py:332-348) -> (B, F, 2C, H, W) fp32."""
        if self.w is None:
            self.prepare()
        m, ops, W = self.m, self.ops, self.w
        B = x.shape[0]
        D, Fr, N, H = m.hidden_size, m.num_frames, m.x_embedder.num_patches, m.num_heads
        T, rpb = B * Fr * N, Fr * N
        dev = x.device
        sc = ops.to_operand(torch.nn.functional.silu(c.float()).contiguous())            # adaLN_modulation[0], final too
        mod = ops.linear(sc, W["ada_w"], W["ada_b"]).float()                               # (B, depth*6D + 2D)
        S = {"B": B, "c": c, "sc": sc, "mod": mod, "blocks": []}

        xp = torch.zeros(T, 64, dtype=torch.float32, device=dev)
        xp[:, : self.kp] = self._patchify(x.float())
        xp = ops.to_operand(xp)
        xs = m.pos_embed.detach().float().reshape(1, N, D).expand(B * Fr, N, D).reshape(T, D).contiguous()
        ops.linear_accum(xs, xp, W["patch_w"], W["patch_b"])
        S["xp"] = xp
        temp = m.temp_embed.detach().float().reshape(Fr, D).contiguous()
        for i in range(m.depth):
            mv = mod[:, i * 6 * D:(i + 1) * 6 * D]
            sh1, sc1, g1, sh2, sc2, g2 = (mv[:, k * D:(k + 1) * D] for k in range(6))
            temporal = bool(i % 2)
            wq, wp, w1, w2 = (W[f"{i}.{n}"] for n in _BLOCK_LINEARS)
            h1 = ops.ln_modulate(xs, sh1, sc1, rpb)
            qkv = ops.linear(h1, wq[0], wq[1])
            o = ops.attention(qkv, B, Fr, N, H, temporal)
            m1 = ops.linear(o, wp[0], wp[1])
            # (`ops.gate_residual_ln` does the residual update and the next LayerNorm-modulate in one pass; built and tested, but
            # timed alone it is slower than the two tuned passes it replaces (101 vs 45 + 40 us), so it is off by default;
            # LATTE_B200_FUSED_RESIDUAL_LN=1 switches it on for this half of the block for A/B runs)
            if _FUSED_RESIDUAL_LN:
                xm, h2 = ops.gate_residual_ln(xs, m1, g1, sh2, sc2, rpb)
            else:
                xm = ops.gate_residual(xs, m1, g1, rpb)
                h2 = ops.ln_modulate(xm, sh2, sc2, rpb)
            u, a = ops.linear_gelu_both(h2, w1[0], w1[1])
            m2 = ops.linear(a, w2[0], w2[1])
            xo = ops.gate_residual(xm, m2, g2, rpb, row_add=temp if i == 0 else None, tokens=N)
            S["blocks"].append((xs, h1, qkv, o, m1, xm, h2, u, a, m2))
            xs = xo
        base = m.depth * 6 * D
        shf, scf = mod[:, base:base + D], mod[:, base + D:base + 2 * D]
        hf = ops.ln_modulate(xs, shf, scf, rpb)
        tok = torch.zeros(T, self.nf, dtype=torch.float32, device=dev)
        ops.linear_accum(tok, hf, W["final_w"], W["final_b"])
        S["x_last"], S["hf"] = xs, hf
        self.saved = S
        return self._unpatchify(tok, B)

    # ---------------------------------------------------------------------------------------------------------------
    def backward(self, dout):
        """dout (B, F, 2C, H, W) -> (grads: {parameter name: fp32 tensor}, dc (B, D) fp32).  Frees the saved activations.
        Bias gradients, the per-sample adaLN gradients (dmod) and every weight gradient are views of buffers zeroed once here;
        the kernels accumulate into them (wgrad through the GEMM's fp32 residual epilogue, reductions with atomics)."""
        m, ops, W, S = self.m, self.ops, self.w, self.saved
        self.saved = None
        B = S["B"]
        D, Fr, N, H, Hm = m.hidden_size, m.num_frames, m.x_embedder.num_patches, m.num_heads, m.mlp_hidden
        T, rpb = B * Fr * N, Fr * N
        dev = dout.device
        mod = S["mod"]
        G = {}
        dmod = torch.zeros_like(mod)
        # all bias gradients in one zeroed buffer: per block [qkv 3D | proj D | fc1 Hm | fc2 D], then final (nf), patch (D)
        per_blk = 3 * D + D + Hm + D
        bias_flat = torch.zeros(m.depth * per_blk + self.nf + D, dtype=torch.float32, device=dev)

        def bias_view(i, off, n):
            return bias_flat[i * per_blk + off: i * per_blk + off + n]

        def wgrad(dy, x):
            g = torch.zeros(dy.shape[1], x.shape[1], dtype=torch.float32, device=dev)
            return ops.wgrad(g, dy, x)

        # ---- final layer (latte.py:197-201) ----
        dtok = self._patchify_out(dout.float())                                         # (T, nf) fp32
        G["final_layer.linear.bias"] = ops.colsum(dtok, bias_flat[m.depth * per_blk: m.depth * per_blk + self.nf])
        dtok16 = ops.to_operand(dtok)
        G["final_layer.linear.weight"] = wgrad(dtok16, S["hf"])
        dtp = torch.zeros(T, 64, dtype=torch.float32, device=dev)
        dtp[:, : self.nf] = dtok
        dhf = ops.dgrad(ops.to_operand(dtp), W["final_wk"])
        dx = torch.zeros(T, D, dtype=torch.float32, device=dev)
        base = m.depth * 6 * D
        ops.ln_modulate_bwd(dhf, S["x_last"], mod[:, base:base + D], mod[:, base + D:base + 2 * D], rpb, dx,
                            dmod[:, base:base + D], dmod[:, base + D:base + 2 * D])
        del dhf, dtp, dtok16

        # ---- blocks, last to first (latte.py:177-181) ----
        for i in reversed(range(m.depth)):
            xs, h1, qkv, o, m1, xm, h2, u, a, m2 = S["blocks"].pop()
            mv = mod[:, i * 6 * D:(i + 1) * 6 * D]
            sh1, sc1, g1, sh2, sc2, g2 = (mv[:, k * D:(k + 1) * D] for k in range(6))
            dv = dmod[:, i * 6 * D:(i + 1) * 6 * D]
            dsh1, dsc1, dg1, dsh2, dsc2, dg2 = (dv[:, k * D:(k + 1) * D] for k in range(6))
            temporal = bool(i % 2)
            wq, wp, w1, w2 = (W[f"{i}.{n}"] for n in _BLOCK_LINEARS)
            p = f"blocks.{i}."
            G[p + "attn.qkv.bias"], G[p + "attn.proj.bias"] = bias_view(i, 0, 3 * D), bias_view(i, 3 * D, D)
            G[p + "mlp.fc1.bias"], G[p + "mlp.fc2.bias"] = bias_view(i, 4 * D, Hm), bias_view(i, 4 * D + Hm, D)
            # x_out = x_mid + g2 * fc2(gelu(fc1(LNmod(x_mid))))
            dm2 = ops.gate_bwd(dx, m2, g2, rpb, dg2, G[p + "mlp.fc2.bias"])
            G[p + "mlp.fc2.weight"] = wgrad(dm2, a)
            del a
            # gelu'(u) stays a separate pass: as an epilogue of this dgrad (`gelu_u=u`, built and tested) it made the N = 4608
            # epilogue the bottleneck of the GEMM -- 327 us against 153 + 129 us for dgrad + gelu_bwd (which also sums the bias gradient)
            da = ops.dgrad(dm2, w2[0])
            du = ops.gelu_bwd(da, u, G[p + "mlp.fc1.bias"])
            del da, dm2
            G[p + "mlp.fc1.weight"] = wgrad(du, h2)
            dh2 = ops.dgrad(du, w1[0])
            del du
            ops.ln_modulate_bwd(dh2, xm, sh2, sc2, rpb, dx, dsh2, dsc2)
            del dh2
            # x_mid = x_in + g1 * proj(attn(qkv(LNmod(x_in))))
            dm1 = ops.gate_bwd(dx, m1, g1, rpb, dg1, G[p + "attn.proj.bias"])
            G[p + "attn.proj.weight"] = wgrad(dm1, o)
            do = ops.dgrad(dm1, wp[0])
            del dm1
            dqkv = ops.attention_bwd(qkv, o, do, B, Fr, N, H, temporal)
            del do
            ops.colsum(dqkv, G[p + "attn.qkv.bias"])
            G[p + "attn.qkv.weight"] = wgrad(dqkv, h1)
            dh1 = ops.dgrad(dqkv, wq[0])
            del dqkv
            ops.ln_modulate_bwd(dh1, xs, sh1, sc1, rpb, dx, dsh1, dsc1)
            del dh1, xs, h1, qkv, o, m1, xm, h2, u, m2

        # ---- patch embedding (latte.py:330-331; pos_embed / temp_embed are frozen, :246-247) ----
        G["x_embedder.proj.bias"] = ops.colsum(dx, bias_flat[m.depth * per_blk + self.nf:])
        gpe = wgrad(ops.to_operand(dx), S["xp"])
        G["x_embedder.proj.weight"] = gpe[:, : self.kp].reshape(m.x_embedder.proj.weight.shape).contiguous()

        # ---- adaLN_modulation of every block + final layer: mod = Linear(SiLU(c)) (latte.py:160-163, 192-195) ----
        dW = ops.ada_outer(dmod, S["sc"])                                               # (depth*6D + 2D, D)
        db = dmod.sum(0)
        for i in range(m.depth):
            G[f"blocks.{i}.adaLN_modulation.1.weight"] = dW[i * 6 * D:(i + 1) * 6 * D]
            G[f"blocks.{i}.adaLN_modulation.1.bias"] = db[i * 6 * D:(i + 1) * 6 * D]
        G["final_layer.adaLN_modulation.1.weight"] = dW[base:base + 2 * D]
        G["final_layer.adaLN_modulation.1.bias"] = db[base:base + 2 * D]
        dsc = ops.ada_dsc(dmod, W["ada_w"])
        c = S["c"].float()
        sg = torch.sigmoid(c)
        dc = dsc * (sg * (1 + c * (1 - sg)))                                            # d silu
        return G, dc


def trainable_names(model):
    """Names, in a fixed order, of the parameters the engine produces gradients for (everything except the embedders that
    feed `c`, whose few-kilobyte graph stays on torch autograd, and the frozen sin-cos tables)."""
    names = ["x_embedder.proj.weight", "x_embedder.proj.bias"]
    for i in range(model.depth):
        for n in _BLOCK_LINEARS:
            names += [f"blocks.{i}.{n}.weight", f"blocks.{i}.{n}.bias"]
        names += [f"blocks.{i}.adaLN_modulation.1.weight", f"blocks.{i}.adaLN_modulation.1.bias"]
    names += ["final_layer.linear.weight", "final_layer.linear.bias",
              "final_layer.adaLN_modulation.1.weight", "final_layer.adaLN_modulation.1.bias"]
    return names


class _LatteTrainFn(torch.autograd.Function):
    """Autograd boundary: inputs (x, c, *parameters) -> output; backward hands each parameter its gradient, so optimizers,
    `clip_grad_norm_`, gradient accumulation and DistributedDataParallel's bucketed all-reduce hooks see ordinary `.grad`s."""

    @staticmethod
    def forward(ctx, engine, names, x, c, *params):
        ctx.engine, ctx.names = engine, names
        ctx.dtypes = [p.dtype for p in params]
        with torch.no_grad():
            engine.prepare()
            out = engine.forward(x.detach(), c.detach())
        return out

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.engine
        if eng.saved is None:
            raise RuntimeError("latte_b200: backward called twice on one training forward (activations are freed after the first)")
        with torch.no_grad():
            G, dc = eng.backward(dout.contiguous())
        grads = tuple(G[n].to(dt) if G[n].dtype != dt else G[n] for n, dt in zip(ctx.names, ctx.dtypes))
        return (None, None, None, dc) + grads


def train_forward(model, ops, dtype, x, c):
    """Forward of one training step with the backward attached.  c = t_embedder(t) + y_embedder(y), computed by the caller with
    torch autograd (a (B, D) graph)."""
    eng = TrainEngine(model, ops, dtype)
    names = trainable_names(model)
    named = dict(model.named_parameters())
    params = [named[n] for n in names]
    return _LatteTrainFn.apply(eng, names, x, c, *params)


def conditioning(model, t, y):
    """c = t_embedder(t) [+ y_embedder(y)] with torch autograd (latte.py:98-123 sincos + MLP, :148-153 table lookup): a
    (B, D) graph of a handful of kernels whose parameters get their gradients from `dc`."""
    import math
    half = model.t_embedder.frequency_embedding_size // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    c = model.t_embedder.mlp(emb.to(model.t_embedder.mlp[0].weight.dtype)).float()
    if model.extras == 2:
        c = c + model.y_embedder.embedding_table(y).float()
    return c
