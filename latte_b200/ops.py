"""Thin torch-tensor wrappers over the per-op C-ABI entry points (used by the parity tests and by
anyone who wants a single fused op).  Device pointers + current stream in, nothing else."""
from __future__ import annotations

import torch

from . import _lib


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _lib.FP16
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"expected a float16/bfloat16 tensor, got {t.dtype}")


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


_SK_FLAGS = {}


def _sk_flags(device):
    """Per-device stream-K ordering flags for `b200_linear` (zeroed once; every launch leaves them zero, include/latte_b200.h)."""
    buf = _SK_FLAGS.get(device)
    if buf is None:
        buf = _SK_FLAGS[device] = torch.zeros(_lib.GEMM_SK_FLAGS, dtype=torch.int64, device=device)
    return buf


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("latte_b200.ops run on CUDA (sm_100a) only; there is no CPU fallback")


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, gelu: bool = False,
           block_n: int = 0) -> torch.Tensor:
    """out16 = (gelu_tanh)(a @ w.T + bias); a [M,K], w [N,K] 16-bit, bias fp32."""
    _need_cuda(a, w, bias)
    assert a.dtype == w.dtype and a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=a.dtype, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().b200_linear(a.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, M, N, K,
                                     _dt(a), _lib.EPI_BIAS_GELU if gelu else _lib.EPI_BIAS, out.data_ptr(), None, None,
                                     0, 1, block_n, None, _stream(a))
    _lib.check(rc, "b200_linear")
    return out


def linear_gate_residual_(resid: torch.Tensor, a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None,
                          gate: torch.Tensor, rows_per_batch: int, block_n: int = 0, stream_k: bool = True) -> torch.Tensor:
    """resid (fp32 [M,N], in place) += gate[row // rows_per_batch] * (a @ w.T + bias); gate fp32 [B, N].
    stream_k=False withholds the flag buffer, which keeps the data-parallel schedule (one add per element)."""
    _need_cuda(resid, a, w, bias, gate)
    assert resid.dtype == torch.float32 and gate.dtype == torch.float32 and resid.is_contiguous()
    assert gate.stride(-1) == 1
    M, K = a.shape
    N = w.shape[0]
    with torch.cuda.device(a.device):
        rc = _lib.load().b200_linear(a.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, M, N, K,
                                     _dt(a), _lib.EPI_GATE_RESIDUAL, None, resid.data_ptr(), gate.data_ptr(),
                                     gate.stride(0), rows_per_batch, block_n,
                                     _sk_flags(a.device).data_ptr() if stream_k else None, _stream(a))
    _lib.check(rc, "b200_linear")
    return resid


def attention(qkv: torch.Tensor, batch: int, frames: int, tokens: int, heads: int, temporal: bool) -> torch.Tensor:
    """qkv [batch*frames*tokens, 3*heads*hd] 16-bit -> out [T, heads*hd] 16-bit."""
    _need_cuda(qkv)
    assert qkv.is_contiguous() and qkv.shape[0] == batch * frames * tokens
    D = qkv.shape[1] // 3
    out = torch.empty(qkv.shape[0], D, dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = _lib.load().b200_attention(qkv.data_ptr(), out.data_ptr(), batch, frames, tokens, heads, D // heads,
                                        _dt(qkv), int(temporal), _stream(qkv))
    _lib.check(rc, "b200_attention")
    return out


def ln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, rows_per_batch: int,
                dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """x fp32 [rows, D]; shift/scale fp32 [B, D] (row stride arbitrary) -> 16-bit [rows, D]."""
    _need_cuda(x, shift, scale)
    assert x.dtype == torch.float32 and x.is_contiguous() and shift.stride(0) == scale.stride(0)
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().b200_ln_modulate(x.data_ptr(), shift.data_ptr(), scale.data_ptr(), shift.stride(0),
                                          rows_per_batch, out.data_ptr(), x.shape[0], x.shape[1], _dt(out), _stream(x))
    _lib.check(rc, "b200_ln_modulate")
    return out


def cross_attention(q: torch.Tensor, kv: torch.Tensor, batch: int, q_rows_per_batch: int, kv_len: int, heads: int,
                    key_bias: torch.Tensor | None = None) -> torch.Tensor:
    """q [batch*q_rows, heads*hd] 16-bit; kv [batch*kv_len, 2*heads*hd] 16-bit ([k | v]); kv_len <= 128 -> out like q.
    key_bias: optional fp32 [batch, 128] additive score bias per key (columns >= kv_len ignored)."""
    _need_cuda(q, kv, key_bias)
    assert q.is_contiguous() and kv.is_contiguous() and q.dtype == kv.dtype
    if key_bias is not None:
        assert key_bias.dtype == torch.float32 and key_bias.is_contiguous() and tuple(key_bias.shape) == (batch, 128)
    D = q.shape[1]
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        rc = _lib.load().b200_cross_attention(q.data_ptr(), kv.data_ptr(), key_bias.data_ptr() if key_bias is not None else None,
                                              out.data_ptr(), batch, q_rows_per_batch, kv_len,
                                              q.shape[1], kv.shape[1], heads, D // heads, _dt(q), _stream(q))
    _lib.check(rc, "b200_cross_attention")
    return out


_U8_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def frames_to_uint8(video: torch.Tensor, mode: str = "pipeline") -> torch.Tensor:
    """Decoded frames [n, c, h, w] in [-1, 1] -> uint8 [n, h, w, c] on the device, byte-identical to the reference expression:
    mode "pipeline" = pipeline_latte.py:775,796 (truncating), mode "sample" = sample.py:122 / sample_ddp.py:172 (+0.5 rounding)."""
    _need_cuda(video)
    assert video.dim() == 4 and video.is_contiguous() and video.dtype in _U8_DT
    n, c, h, w = video.shape
    out = torch.empty(n, h, w, c, dtype=torch.uint8, device=video.device)
    with torch.cuda.device(video.device):
        rc = _lib.load().b200_frames_to_uint8(video.data_ptr(), _U8_DT[video.dtype], n, c, h, w, {"pipeline": 0, "sample": 1}[mode],
                                              out.data_ptr(), _stream(video))
    _lib.check(rc, "b200_frames_to_uint8")
    return out
