"""Thin torch-tensor wrappers over the operator-level C-ABI entry points (vla_op_*).  Torch is only the
owner of device memory and of the current stream; all arithmetic happens in libvla_b200.so."""
from __future__ import annotations

import torch

from . import _lib

ACT = {"none": 0, "gelu": 1, "relu": 2, "swiglu": 3}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vla_adapter_b200 ops need CUDA tensors (there is no CPU fallback)")


def linear(a: torch.Tensor, w: torch.Tensor, bias=None, act="none", colscale=None, resid=None, out=None,
           force_bn: int = 0) -> torch.Tensor:
    """out = epi(a @ w.T). a: (M, K) bf16 (row stride may exceed K), w: (N, K) bf16, bias/colscale fp32."""
    _need_cuda(a, w, bias, colscale, resid, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if act == "swiglu" else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=a.device)
    lib = _lib.load()
    rc = lib.vla_op_gemm(_ptr(a), 0, a.stride(0), M, 1, _ptr(w), w.stride(0), N, K, _ptr(out), 0, out.stride(0),
                         _ptr(bias), _ptr(colscale), _ptr(resid), 0, resid.stride(0) if resid is not None else 0,
                         ACT[act], force_bn, _stream())
    _lib.check(rc)
    return out


def linear_batched(a: torch.Tensor, row0: int, rows: int, w: torch.Tensor, out: torch.Tensor, out_row0: int,
                   bias=None, act="none") -> torch.Tensor:
    """3-D view form: a is (B, R, K); reads rows [row0, row0+rows) of every slab, writes rows
    [out_row0, out_row0+rows) of every slab of out (B, R2, N)."""
    _need_cuda(a, w, out)
    B, R, K = a.shape
    N = w.shape[0]
    lib = _lib.load()
    a_v = a[:, row0:, :]
    o_v = out[:, out_row0:, :]
    rc = lib.vla_op_gemm(a_v.data_ptr(), a.stride(0), a.stride(1), rows, B, _ptr(w), w.stride(0), N, K,
                         o_v.data_ptr(), out.stride(0), out.stride(1), _ptr(bias), None, None, 0, 0, ACT[act], 0,
                         _stream())
    _lib.check(rc)
    return out


def layernorm(x, w, b, eps=1e-6):
    _need_cuda(x, w, b)
    y = torch.empty_like(x)
    rows, dim = x.shape
    _lib.check(_lib.load().vla_op_layernorm(_ptr(x), rows, dim, x.stride(0), _ptr(w), _ptr(b), eps, _ptr(y),
                                             y.stride(0), _stream()))
    return y


def rmsnorm(x, w, eps=1e-6):
    _need_cuda(x, w)
    y = torch.empty_like(x)
    rows, dim = x.shape
    _lib.check(_lib.load().vla_op_rmsnorm(_ptr(x), rows, dim, x.stride(0), _ptr(w), eps, _ptr(y), y.stride(0),
                                           _stream()))
    return y


def norm_linear(x, norm_w, norm_b, W, bias, eps=1e-6, act="none"):
    """act(Linear(Norm(x))) the way the engine runs it: the norm's weight / bias are folded into (copies of) W / bias
    and the GEMM consumes the raw rows; norm_b=None selects RMSNorm.  Returns (rows, N), or (rows, N/2) for swiglu."""
    _need_cuda(x, norm_w, W)
    rows, K = x.shape
    N = W.shape[0]
    rms = norm_b is None
    Wf = W.clone()
    bf = bias.clone() if bias is not None else (None if rms else torch.zeros(N, dtype=torch.float32, device=x.device))
    colsum = torch.empty(N, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.vla_op_fold_norm(_ptr(Wf), N, K, Wf.stride(0), _ptr(norm_w), None if rms else _ptr(norm_b),
                                    None if bf is None else _ptr(bf), _ptr(colsum), _stream()))
    code = ACT[act]
    out = torch.empty((rows, N // 2 if act == "swiglu" else N), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
    _lib.check(lib.vla_op_norm_gemm(_ptr(x), rows, x.stride(0), _ptr(Wf), Wf.stride(0), N, K, _ptr(out), out.stride(0),
                                    None if bf is None else _ptr(bf), _ptr(colsum), int(rms), eps, code, _ptr(stats),
                                    _stream()))
    return out


def block_tail(a, W1, bias1, colscale1, x, norm_w, norm_b, W2, bias2, eps=1e-6, act="none"):
    """x += colscale1 * (a @ W1.T + bias1) in place, then act(Linear(Norm(x))) with the row statistics taken from the
    first GEMM's epilogue (no statistics kernel).  Returns (out, partials); norm_b=None selects RMSNorm."""
    _need_cuda(a, W1, x, norm_w, W2)
    rows, K1 = a.shape
    D, N2 = x.shape[1], W2.shape[0]
    rms = norm_b is None
    Wf = W2.clone()
    bf = bias2.clone() if bias2 is not None else (None if rms else torch.zeros(N2, dtype=torch.float32, device=x.device))
    colsum = torch.empty(N2, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.vla_op_fold_norm(_ptr(Wf), N2, D, Wf.stride(0), _ptr(norm_w), None if rms else _ptr(norm_b),
                                    None if bf is None else _ptr(bf), _ptr(colsum), _stream()))
    out = torch.empty((rows, N2 // 2 if act == "swiglu" else N2), dtype=torch.bfloat16, device=x.device)
    partials = torch.full((rows, 12, 2), float("nan"), dtype=torch.float32, device=x.device)
    _lib.check(lib.vla_op_block_tail(_ptr(a), a.stride(0), rows, _ptr(W1), W1.stride(0), K1, _ptr(x), D, _ptr(bias1),
                                     _ptr(colscale1), _ptr(Wf), Wf.stride(0), N2, _ptr(out), out.stride(0),
                                     None if bf is None else _ptr(bf), _ptr(colsum), int(rms), eps, ACT[act],
                                     _ptr(partials), _stream()))
    return out, partials


def center_crop_u8(images: torch.Tensor, crop_scale: float = 0.9, out_size: int = 224) -> torch.Tensor:
    """(n, H, W, 3) uint8 CUDA -> (n, out, out, 3) uint8: the reference's centre crop (openvla_utils.py:616-648)."""
    _need_cuda(images)
    assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3 and images.is_contiguous()
    n, H, W, _ = images.shape
    out = torch.empty((n, out_size, out_size, 3), dtype=torch.uint8, device=images.device)
    _lib.check(_lib.load().vla_op_center_crop_u8(_ptr(images), _ptr(out), n, H, W, out_size, float(crop_scale), _stream()))
    return out


def attention(qkv, B, S, n_heads, n_kv_heads, hd, causal):
    """qkv: (B*S, (n_heads + 2*n_kv_heads) * hd) packed [q | k | v]."""
    _need_cuda(qkv)
    out = torch.empty((B * S, n_heads * hd), dtype=torch.bfloat16, device=qkv.device)
    q_off, k_off, v_off = 0, n_heads * hd, (n_heads + n_kv_heads) * hd
    _lib.check(_lib.load().vla_op_attention(_ptr(qkv), qkv.stride(0), q_off, k_off, v_off, B, S, n_heads,
                                             n_heads // n_kv_heads, hd, int(causal), _ptr(out), out.stride(0),
                                             _stream()))
    return out


def linear_rope(a, w, bias, cos_t, sin_t, rope_cols, S):
    """out = rope(a @ w.T + bias): q/k/v projection with RoPE fused into the GEMM epilogue (cos_t/sin_t: (S, 32) fp32)."""
    _need_cuda(a, w, bias, cos_t, sin_t)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _lib.check(_lib.load().vla_op_gemm_rope(_ptr(a), a.stride(0), M, _ptr(w), w.stride(0), N, K, _ptr(out), out.stride(0),
                                             _ptr(bias), _ptr(cos_t), _ptr(sin_t), rope_cols, S, _stream()))
    return out


def cross_attention(q, kv, B, Sq, Skv, n_heads, n_kv_heads, hd):
    """q: (B*Sq, n_heads*hd); kv: (B*Skv, 2*n_kv_heads*hd) packed [k | v]; non-causal."""
    _need_cuda(q, kv)
    out = torch.empty((B * Sq, n_heads * hd), dtype=torch.bfloat16, device=q.device)
    k, v = kv, kv[:, n_kv_heads * hd:]
    _lib.check(_lib.load().vla_op_cross_attention(_ptr(q), q.stride(0), Sq, _ptr(k), v.data_ptr(), kv.stride(0), Skv, B,
                                                   n_heads, n_heads // n_kv_heads, hd, 0, _ptr(out), out.stride(0),
                                                   _stream()))
    return out


def set_attention_impl(impl: int) -> None:
    """0 = auto, 1 = mma.sync kernel only, 2 = tcgen05 kernel whenever the head dim allows."""
    _lib.check(_lib.load().vla_set_attention_impl(int(impl)))


def rope_(x, off, n_heads, B, S, theta):
    _need_cuda(x)
    _lib.check(_lib.load().vla_op_rope(_ptr(x), x.stride(0), off, n_heads, B, S, float(theta), _stream()))
    return x
