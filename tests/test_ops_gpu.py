"""Operator-level parity: each CUDA kernel (through the C ABI) against a plain PyTorch fp32 reference of
the same op on the same seeded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _randn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).cuda()


GEMM_SHAPES = [
    # M, N, K
    (128, 256, 64),
    (128, 128, 128),
    (261, 1024, 1024),     # DINOv2 proj at bs=1 (M tail)
    (512, 4304, 1152),     # SigLIP fc1 (N tail: 4304 = 16*256 + 208)
    (512, 1152, 4304),     # SigLIP fc2 (K tail: 4304 = 67*64 + 16)
    (256, 1024, 592),      # patch embed (K tail)
    (1250, 1152, 896),     # Qwen qkv, 2 samples
    (1000, 896, 4864),     # Qwen down
    (8, 2688, 896),        # policy q/k/v at bs=1 (tiny M)
    (3000, 8704, 2176),    # projector fc1
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 224, 256])
def test_gemm_plain(M, N, K, bn):
    from vla_adapter_b200 import ops

    a = _randn(M, K, seed=1)
    w = _randn(N, K, scale=K ** -0.5, seed=2)
    out = ops.linear(a, w, force_bn=bn)
    ref = a.float() @ w.float().T
    torch.cuda.synchronize()
    assert _rel(out, ref) < 5e-3, (M, N, K, bn)
    # bf16 rounding of an fp32-accumulated result: elementwise within 1 bf16 ulp of the fp32 reference
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("act,N,bn", [("none", 1152, 0), ("gelu", 1152, 0), ("relu", 1152, 0), ("none", 896, 224),
                                      ("gelu", 896, 224), ("relu", 1000, 224)])
def test_gemm_epilogue(act, N, bn):
    from vla_adapter_b200 import ops

    M, K = 700, 896
    a = _randn(M, K, seed=3)
    w = _randn(N, K, scale=K ** -0.5, seed=4)
    bias = torch.randn(N, device="cuda")
    ls = torch.rand(N, device="cuda") + 0.5
    resid = _randn(M, N, seed=5)
    out = ops.linear(a, w, bias=bias, act=act, colscale=ls, resid=resid, force_bn=bn)
    v = a.float() @ w.float().T + bias
    if act == "gelu":
        v = torch.nn.functional.gelu(v)
    elif act == "relu":
        v = torch.relu(v)
    ref = resid.float() + ls * v
    assert _rel(out, ref) < 5e-3
    # in-place residual (C aliases resid)
    x = resid.clone()
    ops.linear(a, w, bias=bias, act=act, colscale=ls, resid=x, out=x, force_bn=bn)
    # in place the epilogue rounds to bf16 and then TMA-reduce-adds into C (two roundings, like the reference's bf16
    # `x + y`); out of place it adds the residual in fp32 before the single rounding: equal up to one bf16 ulp
    assert _rel(x, ref) < 5e-3
    assert (x.float() - out.float()).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


def test_gemm_swiglu():
    from vla_adapter_b200 import ops

    M, I, K = 900, 4864, 896
    a = _randn(M, K, seed=6)
    wg = _randn(I, K, scale=K ** -0.5, seed=7)
    wu = _randn(I, K, scale=K ** -0.5, seed=8)
    # interleave rows in groups of 16: g0..15, u0..15, g16..31, u16..31, ...
    w = torch.stack([wg.view(I // 16, 16, K), wu.view(I // 16, 16, K)], dim=1).reshape(2 * I, K).contiguous()
    out = ops.linear(a, w, act="swiglu")
    ref = torch.nn.functional.silu(a.float() @ wg.float().T) * (a.float() @ wu.float().T)
    assert out.shape == (M, I)
    assert _rel(out, ref) < 5e-3


def test_gemm_batched_view():
    from vla_adapter_b200 import ops

    B, R, K, N = 3, 625, 896, 1792
    a = _randn(B, R, K, seed=9)
    w = _randn(N, K, scale=K ** -0.5, seed=10)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(B, 600, N, dtype=torch.bfloat16, device="cuda")
    ops.linear_batched(a, 559, 64, w, out, 3, bias=bias)          # h_a-style slice: rows [559, 623)
    ref = a[:, 559:623].float() @ w.float().T + bias
    assert _rel(out[:, 3:67], ref) < 5e-3
    assert out[:, :3].abs().sum().item() == 0 and out[:, 67:].abs().sum().item() == 0   # nothing else written
    out.zero_()
    ops.linear_batched(a, 0, 512, w, out, 65, bias=bias)           # h_t-style slice: rows [0, 512)
    ref = a[:, :512].float() @ w.float().T + bias
    assert _rel(out[:, 65:577], ref) < 5e-3
    assert out[:, :65].abs().sum().item() == 0 and out[:, 577:].abs().sum().item() == 0


@pytest.mark.parametrize("rows,B", [(8, 5), (8, 64), (64, 3), (1, 7), (16, 9), (25, 4), (32, 6)])
def test_gemm_short_row_views_packed(rows, B):
    """Short row views (the policy's T rows / 64 ActionQuery rows per sample): 128 / rows samples share one M tile.
    Rows that divide 128 take the packed path, 25 (the larger-chunk preset) the plain one; nothing outside the view
    may be written."""
    from vla_adapter_b200 import ops

    R, R2, K, N = 40, 70, 896, 1792
    a = _randn(B, R, K, seed=51)
    w = _randn(N, K, scale=K ** -0.5, seed=52)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(B, R2, N, dtype=torch.bfloat16, device="cuda")
    rows = min(rows, R - 3)
    ops.linear_batched(a, 3, rows, w, out, 5, bias=bias)
    ref = a[:, 3:3 + rows].float() @ w.float().T + bias
    assert _rel(out[:, 5:5 + rows], ref) < 5e-3
    assert out[:, :5].abs().sum().item() == 0 and out[:, 5 + rows:].abs().sum().item() == 0


@pytest.mark.parametrize("dim", [896, 1024, 1152])
def test_layernorm(dim):
    from vla_adapter_b200 import ops

    x = _randn(777, dim, seed=11) * 3 + 0.5
    w = torch.randn(dim, device="cuda")
    b = torch.randn(dim, device="cuda")
    y = ops.layernorm(x, w, b, 1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), w, b, 1e-6)
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


def test_rmsnorm():
    from vla_adapter_b200 import ops

    dim = 896
    x = _randn(1250, dim, seed=12) * 2
    w = (torch.randn(dim) * 0.1 + 1).to(torch.bfloat16).float().cuda()
    y = ops.rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16).float()
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("rows,dim,N,act", [(777, 1024, 3072, "none"), (1000, 1152, 4304, "gelu"), (261, 1024, 4096, "gelu"),
                                            (33, 1152, 3456, "none")])
def test_layernorm_folded_into_gemm(rows, dim, N, act):
    """norm1 -> qkv / norm2 -> fc1 of the ViT blocks: LayerNorm folded into the GEMM (row statistics in the epilogue,
    W * diag(g), bias + W b) against the fp32 LayerNorm -> Linear, and against the unfolded kernel pair."""
    from vla_adapter_b200 import ops

    x = _randn(rows, dim, seed=41) * 3 + 0.7          # non-zero mean: the mean term of the fold matters
    x[:, 5] += 40.0                                   # an outlier channel, as in real ViT residual streams
    g = (torch.randn(dim, device="cuda") * 0.3 + 1).float()
    b = (torch.randn(dim, device="cuda") * 0.2).float()
    W = _randn(N, dim, scale=dim ** -0.5, seed=42)
    bias = torch.randn(N, device="cuda")
    out = ops.norm_linear(x, g, b, W, bias, 1e-6, act)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), g, b, 1e-6) @ W.float().T + bias
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    two_step = ops.linear(ops.layernorm(x, g, b, 1e-6), W, bias=bias, act=act)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 5e-3
    assert _rel(out, ref) < 1.5 * _rel(two_step, ref) + 1e-4   # no worse than rounding the normalised rows to bf16


@pytest.mark.parametrize("act", ["none", "swiglu"])
def test_rmsnorm_folded_into_gemm(act):
    from vla_adapter_b200 import ops

    rows, dim = 1250, 896
    x = _randn(rows, dim, seed=43) * 2
    x[:, 7] *= 30.0
    g = (torch.randn(dim, device="cuda") * 0.1 + 1).float()
    xf = x.float()
    xn = g * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    if act == "swiglu":
        I = 4864
        wg = _randn(I, dim, scale=dim ** -0.5, seed=44)
        wu = _randn(I, dim, scale=dim ** -0.5, seed=45)
        W = torch.stack([wg.view(I // 16, 16, dim), wu.view(I // 16, 16, dim)], dim=1).reshape(2 * I, dim).contiguous()
        out = ops.norm_linear(x, g, None, W, None, 1e-6, act)
        ref = torch.nn.functional.silu(xn @ wg.float().T) * (xn @ wu.float().T)
    else:
        W = _randn(1152, dim, scale=dim ** -0.5, seed=46)
        bias = torch.randn(1152, device="cuda")
        out = ops.norm_linear(x, g, None, W, bias, 1e-6, act)
        ref = xn @ W.float().T + bias
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 5e-3


def _ref_attention(qkv, B, S, H, HKV, hd, causal):
    q, k, v = qkv.float().split([H * hd, HKV * hd, HKV * hd], dim=-1)
    q = q.view(B, S, H, hd).transpose(1, 2)
    k = k.view(B, S, HKV, hd).transpose(1, 2).repeat_interleave(H // HKV, dim=1)
    v = v.view(B, S, HKV, hd).transpose(1, 2).repeat_interleave(H // HKV, dim=1)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if causal:
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B * S, H * hd)


@pytest.mark.parametrize("B,S,H,HKV,hd,causal", [
    (2, 261, 16, 16, 64, False),   # DINOv2
    (2, 256, 16, 16, 72, False),   # SigLIP
    (2, 625, 14, 2, 64, True),     # Qwen2.5 GQA causal
    (1, 609, 14, 2, 64, False),    # bidirectional LLM mode
    (3, 64, 14, 2, 64, True),
    (1, 1, 16, 16, 72, False),
    (2, 128, 16, 16, 64, False),   # exactly one tile
    (1, 129, 14, 2, 64, True),     # one row into the second tile
    (2, 400, 16, 16, 72, True),
    (2, 272, 12, 12, 64, False),   # 16 rows past two full tiles: tensor-core kernel + key-splitting kernel
    (1, 400, 14, 2, 64, False),    # same with GQA
    (3, 389, 12, 12, 64, False),   # 5 rows past three full tiles
])
@pytest.mark.parametrize("impl", [1, 2])   # 1 = mma.sync kernel, 2 = tcgen05/TMEM kernel
def test_attention(B, S, H, HKV, hd, causal, impl):
    from vla_adapter_b200 import ops

    qkv = _randn(B * S, (H + 2 * HKV) * hd, seed=13)
    ops.set_attention_impl(impl)
    try:
        out = ops.attention(qkv, B, S, H, HKV, hd, causal)
        torch.cuda.synchronize()
    finally:
        ops.set_attention_impl(0)
    ref = _ref_attention(qkv, B, S, H, HKV, hd, causal)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,Sq,Skv,H,HKV,hd", [
    (3, 8, 585, 8, 8, 112),     # Bridge-Attention, LIBERO: 8 queries x (8 + 65 + 512) keys
    (2, 25, 602, 8, 8, 112),    # larger chunk: two 16-row query tiles
    (2, 8, 329, 8, 8, 112),     # one image: 8 + 65 + 256 keys
    (1, 5, 261, 14, 2, 64),     # GQA, ragged key split across the four warps
    (2, 8, 130, 8, 8, 112),     # barely enough keys: the last warps get none
])
@pytest.mark.parametrize("impl", [1, 0])   # 1 = generic mma.sync kernel, 0 = split-KV kernel
def test_cross_attention_few_queries(B, Sq, Skv, H, HKV, hd, impl):
    from vla_adapter_b200 import ops

    q = _randn(B * Sq, H * hd, seed=31)
    kv = _randn(B * Skv, 2 * HKV * hd, seed=32)
    ops.set_attention_impl(impl)
    try:
        out = ops.cross_attention(q, kv, B, Sq, Skv, H, HKV, hd)
        torch.cuda.synchronize()
    finally:
        ops.set_attention_impl(0)
    qf = q.float().view(B, Sq, H, hd).transpose(1, 2)
    kf = kv[:, :HKV * hd].float().view(B, Skv, HKV, hd).transpose(1, 2).repeat_interleave(H // HKV, dim=1)
    vf = kv[:, HKV * hd:].float().view(B, Skv, HKV, hd).transpose(1, 2).repeat_interleave(H // HKV, dim=1)
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) / math.sqrt(hd), dim=-1) @ vf).transpose(1, 2).reshape(B * Sq, H * hd)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert _rel(out, ref) < 1e-2


def test_attention_large_scores_rescale():
    """Scores that grow by far more than 2^8 from one key tile to the next exercise the lazy O rescale of the
    tcgen05 kernel; the softmax is then nearly one-hot and must still match the fp32 reference."""
    from vla_adapter_b200 import ops

    B, S, H, hd = 1, 640, 16, 64
    qkv = _randn(B * S, 3 * H * hd, seed=21)
    k = qkv[:, H * hd:2 * H * hd].view(S, H, hd)
    k *= torch.linspace(0.2, 6.0, S, device="cuda").to(torch.bfloat16)[:, None, None]   # later keys score higher
    ops.set_attention_impl(2)
    try:
        out = ops.attention(qkv, B, S, H, H, hd, False)
        torch.cuda.synchronize()
    finally:
        ops.set_attention_impl(0)
    ref = _ref_attention(qkv, B, S, H, H, hd, False)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 1.5e-2


@pytest.mark.parametrize("B,S", [(2, 625), (1, 300), (3, 129)])
def test_gemm_with_fused_rope_equals_gemm_then_rope(B, S):
    """The Qwen2 q/k/v projection with RoPE in the GEMM epilogue is bit-identical to the plain GEMM followed by the
    stand-alone RoPE kernel (which test_rope pins against the HF arithmetic)."""
    from vla_adapter_b200 import ops

    H, HKV, hd, theta, K = 14, 2, 64, 1e6, 896
    N = (H + 2 * HKV) * hd
    a = _randn(B * S, K, seed=41)
    w = _randn(N, K, scale=K ** -0.5, seed=42)
    bias = torch.randn(N, device="cuda")
    ref = ops.linear(a, w, bias=bias)
    ops.rope_(ref, 0, H + HKV, B, S, theta)
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float64) / hd))
    ang = (torch.arange(S, dtype=torch.float32)[:, None] * inv.float()[None, :]).double()
    cos_t = ang.cos().float().to(torch.bfloat16).float().cuda().contiguous()
    sin_t = ang.sin().float().to(torch.bfloat16).float().cuda().contiguous()
    out = ops.linear_rope(a, w, bias, cos_t, sin_t, (H + HKV) * hd, S)
    torch.cuda.synchronize()
    assert torch.equal(out[:, (H + HKV) * hd:], ref[:, (H + HKV) * hd:])      # v heads untouched
    diff = (out.float() - ref.float()).abs()
    assert diff.max().item() <= 0.0625 and (diff > 0).float().mean().item() < 1e-3   # table rounding flips only


def test_rope():
    from vla_adapter_b200 import ops

    B, S, H, HKV, hd, theta = 2, 625, 14, 2, 64, 1e6
    x = _randn(B * S, (H + 2 * HKV) * hd, seed=14)
    y = x.clone()
    ops.rope_(y, 0, H + HKV, B, S, theta)
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    ang = torch.arange(S).float()[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos().to(torch.bfloat16).cuda(), emb.sin().to(torch.bfloat16).cuda()
    xv = x[:, : (H + HKV) * hd].view(B, S, H + HKV, hd)
    rot = torch.cat([-xv[..., hd // 2:], xv[..., : hd // 2]], -1)
    ref = xv * cos[None, :, None, :] + rot * sin[None, :, None, :]      # bf16 eager arithmetic, like HF
    got = y[:, : (H + HKV) * hd].view(B, S, H + HKV, hd)
    assert (got.float() - ref.float()).abs().max().item() <= 0.0625     # rare 1-ulp table flips only
    assert (got.float() - ref.float()).abs().mean().item() < 1e-3
    assert torch.equal(y[:, (H + HKV) * hd:], x[:, (H + HKV) * hd:])     # v untouched


@pytest.mark.parametrize("rows,K1,D,N2,act,rms,ls", [
    (1000, 1024, 1024, 4096, "gelu", False, True),    # DINOv2: proj (+LayerScale) -> norm2 -> fc1 + GELU
    (777, 4304, 1152, 3456, "none", False, False),    # SigLIP: fc2 -> next block's norm1 -> qkv
    (1250, 896, 896, 9728, "swiglu", True, False),    # Qwen: o-proj -> post_attention RMSNorm -> gate|up + SwiGLU
    (300, 4864, 896, 1152, "none", True, False),      # Qwen: down-proj -> next layer's input RMSNorm -> q|k|v
])
def test_block_tail_statistics_from_the_producing_gemm(rows, K1, D, N2, act, rms, ls):
    """The residual GEMM with the staged (TMA-loaded, fp32-added) residual leaves per-row partial sums that the next,
    norm-folded GEMM finishes in its epilogue: same result as the statistics-kernel path and as fp32 torch."""
    from vla_adapter_b200 import ops

    a = _randn(rows, K1, seed=41)
    W1 = _randn(D, K1, scale=K1 ** -0.5, seed=42)
    b1 = torch.randn(D, device="cuda") * 0.1
    cs1 = (torch.rand(D, device="cuda") + 0.5) if ls else None
    x0 = _randn(rows, D, seed=43) * 2 + 0.7           # a residual stream with a DC offset
    nw = torch.rand(D, device="cuda") + 0.5
    nb = None if rms else torch.randn(D, device="cuda") * 0.1
    W2 = _randn(N2, D, scale=D ** -0.5, seed=44)
    b2 = None if act == "swiglu" else torch.randn(N2, device="cuda") * 0.1
    x = x0.clone()
    out, partials = ops.block_tail(a, W1, b1, cs1, x, nw, nb, W2, b2, eps=1e-6, act=act)
    torch.cuda.synchronize()
    # (1) the in-place residual update: one rounding of the fp32 sum
    upd = a.float() @ W1.float().T + b1
    if ls:
        upd = upd * cs1
    x_ref = (x0.float() + upd)
    assert (x.float() - x_ref).abs().max().item() <= 2 ** -8 * x_ref.abs().max().item() + 1e-3
    # (2) the partial sums add up to the row sums of the new x (slots nobody owns are zero, none is left NaN)
    assert torch.isfinite(partials).all()
    s = partials.sum(1)
    assert torch.allclose(s[:, 0], x_ref.sum(1), rtol=2e-3, atol=0.5)
    assert torch.allclose(s[:, 1], (x_ref * x_ref).sum(1), rtol=2e-3)
    # (3) the norm-folded consumer: against the statistics-kernel path on the same x, and against fp32 torch
    out_k = ops.norm_linear(x, nw, nb, W2, b2, eps=1e-6, act=act)
    xf = x.float()
    if rms:
        xn = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)) * nw
    else:
        xn = torch.nn.functional.layer_norm(xf, (D,), nw, nb, 1e-6)
    y = xn @ W2.float().T
    if b2 is not None:
        y = y + b2
    if act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "swiglu":
        y = y.view(rows, N2 // 32, 2, 16)
        y = (torch.nn.functional.silu(y[:, :, 0]) * y[:, :, 1]).reshape(rows, N2 // 2)
    assert _rel(out, y) < 1e-2
    assert _rel(out, out_k.float()) < 4e-3
