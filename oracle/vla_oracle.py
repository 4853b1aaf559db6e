"""CPU oracle for VLA-Adapter's batched predict_action path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (vla_adapter_b200/, libvla_b200.so) imports or
executes this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may, and there only as the checker / the timed CPU baseline.

It is a plain torch-CPU restatement of the reference algorithm, function by function, each citing the
reference file:line it follows (paths relative to /root/reference):
    MP = prismatic/extern/hf/modeling_prismatic.py      AH = prismatic/models/action_heads.py
    PJ = prismatic/models/projectors.py                 TU = prismatic/training/train_utils.py
    K  = prismatic/vla/constants.py                     FV = prismatic/models/film_vit_wrapper.py
Arithmetic that lives in un-vendored dependencies is restated from their published algorithms:
    timm==0.9.10 VisionTransformer (pyproject.toml:45; the call sites are MP:132-142 and the in-repo
    restatement of timm's _intermediate_layers at FV:124-137,153-168, block formula FV:69,75),
    transformers Qwen2ForCausalLM (pyproject.toml:50; call site MP:834-845).

Pinning (see oracle/make_golden.py and tests/test_oracle_golden.py): the oracle is checked against
outputs of the UNMODIFIED reference modules (MP, AH, PJ, TU, K imported from /root/reference through
oracle/ref_shim.py, with stock transformers-5.5 Qwen2 and a timm stand-in) on seeded weights/inputs, the
vectors being committed under tests/golden/.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), and timm itself is not installable here, so the ViT restatement is additionally
cross-checked against transformers' independent Dinov2WithRegisters / Siglip implementations.

`dtype=torch.float32` gives the fp32 "truth"; `dtype=torch.bfloat16` reproduces the reference's own
precision (every op rounds to bf16, like eager PyTorch on bf16 tensors).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

# ---- constants (K:11-15)
IGNORE_INDEX = -100
ACTION_TOKEN_BEGIN_IDX = 151386
STOP_INDEX = 2
NUM_TOKENS = 64

D_DINO, F_DINO, H_VIT = 1024, 4096, 16
D_SIG, F_SIG = 1152, 4304
D_LLM, I_LLM, HQ, HKV, HD = 896, 4864, 14, 2, 64
HEAD_HEADS = 8


@dataclass
class OracleConfig:
    n_images: int = 2
    chunk_len: int = 8          # NUM_ACTIONS_CHUNK (K:29)
    action_dim: int = 7         # ACTION_DIM
    proprio_dim: int = 8        # PROPRIO_DIM
    pro: bool = False           # use_pro_version (AH:29)
    dino_depth: int = 24
    siglip_depth: int = 27
    llm_layers: int = 24
    vocab_size: int = 151936
    causal: bool = True
    rope_theta: float = 1e6
    rms_eps: float = 1e-6
    norm_stats: dict = field(default_factory=lambda: None)

    @property
    def num_patches(self) -> int:
        return 256 * self.n_images


# =====================================================================================================
# seeded synthetic weights, keyed by the reference state_dict names (prefix vla./head./proprio.)
# =====================================================================================================
def _gen(name: str, seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def make_weights(cfg: OracleConfig, seed: int = 0, dtype=torch.bfloat16) -> dict[str, torch.Tensor]:
    """Random weights drawn per tensor name (order independent).  Degenerate reference inits are avoided on
    purpose: ActionQuery table (zeros at MP:376), gating_factor (zeros at AH:214/320), LayerScale
    (1e-5 in timm) are all drawn non-trivially so that parity tests can see them."""
    W: dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=True, gain=1.0):
        W[name + ".weight"] = (torch.randn(out_f, in_f, generator=_gen(name + ".weight", seed)) * (gain / math.sqrt(in_f))).to(dtype)
        if bias:
            W[name + ".bias"] = (torch.randn(out_f, generator=_gen(name + ".bias", seed)) * 0.05).to(dtype)

    def ln(name, dim, bias=True):
        W[name + ".weight"] = (1.0 + 0.1 * torch.randn(dim, generator=_gen(name + ".weight", seed))).to(dtype)
        if bias:
            W[name + ".bias"] = (0.05 * torch.randn(dim, generator=_gen(name + ".bias", seed))).to(dtype)

    def tower(pfx, D, Fh, depth, dino):
        W[pfx + "patch_embed.proj.weight"] = (torch.randn(D, 3, 14, 14, generator=_gen(pfx + "pe.w", seed)) / math.sqrt(588)).to(dtype)
        W[pfx + "patch_embed.proj.bias"] = (torch.randn(D, generator=_gen(pfx + "pe.b", seed)) * 0.05).to(dtype)
        W[pfx + "pos_embed"] = (torch.randn(1, 256, D, generator=_gen(pfx + "pos", seed)) * 0.2).to(dtype)
        if dino:
            W[pfx + "cls_token"] = (torch.randn(1, 1, D, generator=_gen(pfx + "cls", seed)) * 0.5).to(dtype)
            W[pfx + "reg_token"] = (torch.randn(1, 4, D, generator=_gen(pfx + "reg", seed)) * 0.5).to(dtype)
        for i in range(depth):
            b = f"{pfx}blocks.{i}."
            ln(b + "norm1", D)
            lin(b + "attn.qkv", 3 * D, D)
            lin(b + "attn.proj", D, D)
            ln(b + "norm2", D)
            lin(b + "mlp.fc1", Fh, D)
            lin(b + "mlp.fc2", D, Fh)
            if dino:
                W[b + "ls1.scale_factor"] = (0.2 + 0.3 * torch.rand(D, generator=_gen(b + "ls1", seed))).to(dtype)
                W[b + "ls2.scale_factor"] = (0.2 + 0.3 * torch.rand(D, generator=_gen(b + "ls2", seed))).to(dtype)
        ln(pfx + "norm", D)  # final norm: present in the state dict, NOT on the path (FV:160 norm=False)

    tower("vla.vision_backbone.featurizer.", D_DINO, F_DINO, cfg.dino_depth, True)
    tower("vla.vision_backbone.fused_featurizer.", D_SIG, F_SIG, cfg.siglip_depth, False)
    lin("vla.projector.fc1", 4 * (D_DINO + D_SIG), D_DINO + D_SIG)
    lin("vla.projector.fc2", D_LLM, 4 * (D_DINO + D_SIG))
    lin("vla.projector.fc3", D_LLM, D_LLM)
    lm = "vla.language_model.model."
    W[lm + "embed_tokens.weight"] = (torch.randn(cfg.vocab_size, D_LLM, generator=_gen("embed", seed)) * 0.5).to(dtype)
    W["vla.action_queries.weight"] = (torch.randn(NUM_TOKENS, D_LLM, generator=_gen("aq", seed)) * 0.5).to(dtype)
    for i in range(cfg.llm_layers):
        b = f"{lm}layers.{i}."
        ln(b + "input_layernorm", D_LLM, bias=False)
        lin(b + "self_attn.q_proj", HQ * HD, D_LLM)
        lin(b + "self_attn.k_proj", HKV * HD, D_LLM)
        lin(b + "self_attn.v_proj", HKV * HD, D_LLM)
        lin(b + "self_attn.o_proj", D_LLM, D_LLM, bias=False, gain=0.5)
        ln(b + "post_attention_layernorm", D_LLM, bias=False)
        lin(b + "mlp.gate_proj", I_LLM, D_LLM, bias=False)
        lin(b + "mlp.up_proj", I_LLM, D_LLM, bias=False)
        lin(b + "mlp.down_proj", D_LLM, I_LLM, bias=False, gain=0.5)
    ln(lm + "norm", D_LLM, bias=False)

    hm = "head.model."
    ln(hm + "layer_norm1", cfg.action_dim * D_LLM)
    lin(hm + "fc1", D_LLM, cfg.action_dim * D_LLM)
    for i in range(24):
        b = f"{hm}mlp_resnet_blocks.{i}."
        ln(b + "ffn.0", D_LLM)
        lin(b + "ffn.1", D_LLM, D_LLM)
        names = ["q_proj", "k_self", "v_self", "k_adapter", "v_adapter", "k_task", "v_task", "o_proj"] if cfg.pro else \
                ["q_proj", "k_proj", "v_proj", "o_proj"]
        for nme in names:
            lin(b + nme, D_LLM, D_LLM)
        W[b + "gating_factor"] = torch.randn(1, generator=_gen(b + "g", seed)).to(dtype)
        if cfg.pro:
            lin(b + "film_gen.0", 2 * D_LLM, D_LLM)  # in the state dict, never executed (AH:403-406)
    ln(hm + "layer_norm2", D_LLM)
    lin(hm + "fc2", cfg.action_dim, D_LLM)
    lin("proprio.fc1", D_LLM, cfg.proprio_dim)
    lin("proprio.fc2", D_LLM, D_LLM)
    return W


def make_inputs(cfg: OracleConfig, batch: int, prompt_len: int, seed: int = 0):
    """Synthetic LIBERO-shaped observation batch: uint8 images through the processor's two
    normalisations (preprocessor_config.json means/stds; PP:128-145), random prompt ids, clipped proprio."""
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    img = torch.randint(0, 256, (batch, cfg.n_images, 3, 224, 224), generator=g).float() / 255.0
    m0, s0 = torch.tensor([0.485, 0.456, 0.406]).view(1, 1, 3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(1, 1, 3, 1, 1)
    m1, s1 = torch.full((1, 1, 3, 1, 1), 0.5), torch.full((1, 1, 3, 1, 1), 0.5)
    pix = torch.cat([(img - m0) / s0, (img - m1) / s1], dim=2)          # (B, n, 6, H, W): DINOv2 then SigLIP
    pixel_values = pix.reshape(batch, cfg.n_images * 6, 224, 224).to(torch.bfloat16)   # OU:786 .to(bf16)
    hi = min(cfg.vocab_size, 151643)
    input_ids = torch.randint(3, hi, (batch, prompt_len), generator=g, dtype=torch.int64)
    proprio = torch.randn(batch, cfg.proprio_dim, generator=g).clamp(-1, 1)       # OU:671-701 clips to [-1, 1]
    return pixel_values, input_ids, proprio


# =====================================================================================================
# token / index handling (integer, bit-exact)
# =====================================================================================================
def prepare_inputs(input_ids: torch.Tensor):
    """MP:923-937 + MP:748-784 + TU:8-41, batched.  Returns (ext_ids, labels, all_actions_mask)."""
    B, L = input_ids.shape
    labels = torch.full_like(input_ids, IGNORE_INDEX)                                   # MP:923-924
    placeholder = torch.ones((B, NUM_TOKENS), dtype=input_ids.dtype)                    # MP:751-753
    ext = torch.cat([input_ids, placeholder], dim=-1)                                   # MP:754
    stop = torch.ones((B, 1), dtype=input_ids.dtype) * STOP_INDEX                       # MP:757
    ext = torch.cat([ext, stop], dim=-1)                                                # MP:758
    lab_ext = torch.ones((B, ext.shape[-1] - L), dtype=labels.dtype) * (ACTION_TOKEN_BEGIN_IDX + 1)   # MP:774-778
    labels = torch.cat([labels, lab_ext], dim=-1)                                       # MP:779
    labels[:, -1] = STOP_INDEX                                                          # MP:782
    cumsum = torch.cumsum(labels != IGNORE_INDEX, dim=1)                                # TU:11-14
    # ACTION_DIM in TU is the platform constant; the union of both masks does not depend on it (TU:17, 35)
    action_tokens = labels > ACTION_TOKEN_BEGIN_IDX                                     # TU:20, 38
    mask = action_tokens & (cumsum >= 1)                                                # current | next (MP:460)
    return ext, labels, mask


def aq_index_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """Index form of _replace_input_embeddings (MP:442-452): the k-th True column of a row receives
    ActionQuery row k; other columns get -1."""
    idx = torch.cumsum(mask.to(torch.int32), dim=1) - 1
    return torch.where(mask, idx, torch.full_like(idx, -1)).to(torch.int32)


# =====================================================================================================
# vision towers (timm 0.9.10 VisionTransformer semantics) + projector
# =====================================================================================================
def vit_block(x, W, b, D, dino):
    """timm Block: x + ls1(attn(norm1(x))); x + ls2(mlp(norm2(x)))  (FV:69, 75; LayerScale MP:58-59)."""
    B, N, _ = x.shape
    h = F.layer_norm(x, (D,), W[b + "norm1.weight"], W[b + "norm1.bias"], 1e-6)
    qkv = F.linear(h, W[b + "attn.qkv.weight"], W[b + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, H_VIT, D // H_VIT).permute(2, 0, 3, 1, 4)
    a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])              # timm fused_attn, scale hd^-0.5
    a = a.transpose(1, 2).reshape(B, N, D)
    a = F.linear(a, W[b + "attn.proj.weight"], W[b + "attn.proj.bias"])
    if dino:
        a = a * W[b + "ls1.scale_factor"]
    x = x + a
    h = F.layer_norm(x, (D,), W[b + "norm2.weight"], W[b + "norm2.bias"], 1e-6)
    h = F.linear(h, W[b + "mlp.fc1.weight"], W[b + "mlp.fc1.bias"])
    h = F.gelu(h)                                                           # exact erf GELU
    h = F.linear(h, W[b + "mlp.fc2.weight"], W[b + "mlp.fc2.bias"])
    if dino:
        h = h * W[b + "ls2.scale_factor"]
    return x + h


def vit_tower(img, W, pfx, D, depth, dino):
    """get_intermediate_layers(n={depth-2}) with prefix tokens stripped and no final norm
    (MP:141-142; FV:124-137 patch_embed -> _pos_embed -> blocks, FV:153-168 strip prefix)."""
    x = F.conv2d(img, W[pfx + "patch_embed.proj.weight"], W[pfx + "patch_embed.proj.bias"], stride=14)
    x = x.flatten(2).transpose(1, 2)                                        # (B, 256, D)
    x = x + W[pfx + "pos_embed"]                                            # no_embed_class: pos on patches only
    if dino:
        B = x.shape[0]
        x = torch.cat([W[pfx + "cls_token"].expand(B, -1, -1), W[pfx + "reg_token"].expand(B, -1, -1), x], dim=1)
    for i in range(depth - 1):                                              # blocks 0 .. depth-2; the last block
        x = vit_block(x, W, f"{pfx}blocks.{i}.", D, dino)                   # only produces discarded output
    return x[:, 5:] if dino else x


def vision_backbone(pixel_values, W, cfg: OracleConfig):
    """PrismaticVisionBackbone.forward, multi-image branch (MP:216-237) (== MP:211-214 for one image)."""
    out = []
    for img in torch.split(pixel_values, [6] * cfg.n_images, dim=1):
        a, b = torch.split(img, [3, 3], dim=1)
        p = vit_tower(a, W, "vla.vision_backbone.featurizer.", D_DINO, cfg.dino_depth, True)
        q = vit_tower(b, W, "vla.vision_backbone.fused_featurizer.", D_SIG, cfg.siglip_depth, False)
        out.append(torch.cat([p, q], dim=2))
    return torch.cat(out, dim=1)


def projector(x, W):
    """PrismaticProjector.forward, fused branch (MP:267-271)."""
    x = F.gelu(F.linear(x, W["vla.projector.fc1.weight"], W["vla.projector.fc1.bias"]))
    x = F.gelu(F.linear(x, W["vla.projector.fc2.weight"], W["vla.projector.fc2.bias"]))
    return F.linear(x, W["vla.projector.fc3.weight"], W["vla.projector.fc3.bias"])


# =====================================================================================================
# LLM input assembly + Qwen2 prefill (transformers Qwen2ForCausalLM semantics)
# =====================================================================================================
def assemble(ext_ids, mask, patches, W):
    """embed (MP:936) -> _replace_input_embeddings (MP:418-454) -> _build_multimodal_attention (MP:500-502)."""
    emb = W["vla.language_model.model.embed_tokens.weight"][ext_ids]
    aq = W["vla.action_queries.weight"]
    idx = aq_index_from_mask(mask).long()
    emb = torch.where(mask.unsqueeze(-1), aq[idx.clamp(min=0)], emb)
    return torch.cat([emb[:, :1], patches, emb[:, 1:]], dim=1)


def _rms(x, w, eps):
    """Qwen2RMSNorm.forward: fp32 statistics, cast back, then weight."""
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def _rotate_half(x):
    return torch.cat((-x[..., x.shape[-1] // 2:], x[..., : x.shape[-1] // 2]), dim=-1)


def qwen_prefill(x, W, cfg: OracleConfig):
    """Returns the 25 hidden states of output_hidden_states=True: [0] = inputs_embeds, [i] = output of layer i,
    [-1] = final-RMSNorm(output of the last layer) (transformers tie_last_hidden_states)."""
    B, S, _ = x.shape
    lm = "vla.language_model.model."
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, HD, 2, dtype=torch.int64).float() / HD))
    ang = torch.arange(S).float()[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], dim=-1)
    cos, sin = emb.cos().to(x.dtype)[None, None], emb.sin().to(x.dtype)[None, None]
    states = [x]
    for i in range(cfg.llm_layers):
        b = f"{lm}layers.{i}."
        h = _rms(x, W[b + "input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(h, W[b + "self_attn.q_proj.weight"], W[b + "self_attn.q_proj.bias"]).view(B, S, HQ, HD).transpose(1, 2)
        k = F.linear(h, W[b + "self_attn.k_proj.weight"], W[b + "self_attn.k_proj.bias"]).view(B, S, HKV, HD).transpose(1, 2)
        v = F.linear(h, W[b + "self_attn.v_proj.weight"], W[b + "self_attn.v_proj.bias"]).view(B, S, HKV, HD).transpose(1, 2)
        q = (q * cos) + (_rotate_half(q) * sin)
        k = (k * cos) + (_rotate_half(k) * sin)
        k = k.repeat_interleave(HQ // HKV, dim=1)                           # repeat_kv
        v = v.repeat_interleave(HQ // HKV, dim=1)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=cfg.causal)
        a = a.transpose(1, 2).reshape(B, S, HQ * HD)
        x = x + F.linear(a, W[b + "self_attn.o_proj.weight"])
        h = _rms(x, W[b + "post_attention_layernorm.weight"], cfg.rms_eps)
        h = F.silu(F.linear(h, W[b + "mlp.gate_proj.weight"])) * F.linear(h, W[b + "mlp.up_proj.weight"])
        x = x + F.linear(h, W[b + "mlp.down_proj.weight"])
        states.append(x)
    states[-1] = _rms(x, W[lm + "norm.weight"], cfg.rms_eps)
    return states


def gather_hidden(states, num_patches, prompt_len):
    """MP:848-862 batched the way vla-scripts/finetune.py:398-409 batches it.  Quirks kept on purpose:
    h_t = rows [0, NP) = [tok0, patch0..patch_{NP-2}];  h_a = rows [NP+L-1, NP+L-1+64) = [last prompt token,
    AQ0..AQ62] with L = prompt length and NUM_PROMPT_TOKENS = L-1 (MP:927)."""
    npt = prompt_len - 1
    out = []
    for item in states:
        h_a = item[:, num_patches + npt: num_patches + npt + NUM_TOKENS].unsqueeze(1)
        h_t = item[:, :num_patches].unsqueeze(1)
        out.append(torch.cat([h_t, h_a], dim=2))
    return torch.cat(out, dim=1)                                            # (B, 25, NP+64, D)


# =====================================================================================================
# Bridge-Attention policy head
# =====================================================================================================
def _head_rope_tables(n, dtype):
    """RotaryPositionEmbedding.forward (AH:160-164), hd = 112, AS DEPLOYED: get_action_head casts the whole head
    with `.to(torch.bfloat16)` (experiments/robot/openvla_utils.py:515; training does the same at
    vla-scripts/finetune.py:281), which also casts the non-persistent `inv_freq` buffer (AH:158).  `t` is then
    created in inv_freq's dtype (AH:161), so positions, the outer product and therefore the ANGLES are bf16
    (positions above 256 are not even all representable).  Pinned bit-exactly by tests/golden/libero_pro.npz.
    The fp32 "truth" keeps these deployed angles and evaluates cos/sin on them in fp32."""
    hd = D_LLM // HEAD_HEADS
    inv = (1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))).to(torch.bfloat16)
    t = torch.arange(n, dtype=torch.bfloat16)
    fr = torch.einsum("i,j->ij", t, inv)
    emb = torch.cat([fr, fr], dim=-1)
    if dtype == torch.bfloat16:
        return emb.cos(), emb.sin()
    return emb.float().cos().to(dtype), emb.float().sin().to(dtype)


def _apply_rope_one(x, cos, sin):
    """apply_rope (AH:125-146) for one tensor: interleaved pairs, concat-style table."""
    x1, x2 = x[..., ::2], x[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).reshape_as(x)
    return (x * cos[None, None]) + (rot * sin[None, None])


def head_block(x, h_t, h_a, p, W, b, pro):
    """MLPResNetBlock.forward (AH:218-283) / MLPResNetBlock_Pro.forward (AH:337-410)."""
    B, T, C = x.shape
    hd = C // HEAD_HEADS
    g = torch.tanh(W[b + "gating_factor"])
    cond = torch.cat((h_a, p), dim=1)                                       # AH:233 / AH:347
    lin = lambda n, t: F.linear(t, W[b + n + ".weight"], W[b + n + ".bias"])
    heads = lambda t: t.view(B, t.shape[1], HEAD_HEADS, hd).transpose(1, 2)
    q = heads(lin("q_proj", x))
    if pro:
        ks, vs = heads(lin("k_self", x)), heads(lin("v_self", x))
        kc, vc = heads(lin("k_adapter", cond)), heads(lin("v_adapter", cond))
        kt, vt = heads(lin("k_task", h_t)), heads(lin("v_task", h_t))
        cm, sm = _head_rope_tables(T, x.dtype)
        q, ks = _apply_rope_one(q, cm, sm), _apply_rope_one(ks, cm, sm)
        ca, sa = _head_rope_tables(cond.shape[1], x.dtype)
        kc = _apply_rope_one(kc, ca, sa)
        ct, st = _head_rope_tables(h_t.shape[1], x.dtype)
        kt = _apply_rope_one(kt, ct, st)
    else:
        ks, vs = heads(lin("k_proj", x)), heads(lin("v_proj", x))
        kc, vc = heads(lin("k_proj", cond)), heads(lin("v_proj", cond))
        kt, vt = heads(lin("k_proj", h_t)), heads(lin("v_proj", h_t))
    s = torch.cat([q @ ks.transpose(-2, -1), (q @ kc.transpose(-2, -1)) * 1, (q @ kt.transpose(-2, -1)) * g], dim=-1)
    s = s / math.sqrt(hd)
    w = torch.softmax(s, dim=-1)
    o = w @ torch.cat([vs, vc, vt], dim=2)
    o = o.transpose(1, 2).contiguous().view(B, T, C)
    o = lin("o_proj", o)
    y = F.layer_norm(o + x, (C,), W[b + "ffn.0.weight"], W[b + "ffn.0.bias"], 1e-5)   # AH:281 / AH:409: no outer residual
    return F.relu(F.linear(y, W[b + "ffn.1.weight"], W[b + "ffn.1.bias"]))


def policy_head(multi, proprio, W, cfg: OracleConfig, num_task_tokens, dtype, taps=None):
    """L1RegressionActionHead.predict_action (AH:43-81) + MLPResNet.forward (AH:111-121) + ProprioProjector
    (PJ:19-24).  `dtype` is bf16 for the faithful run (hard cast at AH:53) or fp32 for the truth run."""
    B = multi.shape[0]
    hm = "head.model."
    pr = proprio.reshape(B, -1).to(dtype)
    pf = F.linear(F.gelu(F.linear(pr, W["proprio.fc1.weight"], W["proprio.fc1.bias"])), W["proprio.fc2.weight"], W["proprio.fc2.bias"])
    pf = pf.unsqueeze(1)
    h_t_all, h_a_all = multi[:, :, :num_task_tokens], multi[:, :, num_task_tokens:]
    x = torch.zeros((B, cfg.action_dim * cfg.chunk_len, D_LLM), dtype=dtype).reshape(B, cfg.chunk_len, -1)
    x = F.layer_norm(x, (x.shape[-1],), W[hm + "layer_norm1.weight"], W[hm + "layer_norm1.bias"], 1e-5)
    x = F.relu(F.linear(x, W[hm + "fc1.weight"], W[hm + "fc1.bias"]))
    if taps is not None:
        taps["head_x.0"] = x
    for i in range(24):
        x = head_block(x, h_t_all[:, i + 1], h_a_all[:, i + 1], pf, W, f"{hm}mlp_resnet_blocks.{i}.", cfg.pro)
        if taps is not None:
            taps[f"head_x.{i + 1}"] = x
    x = F.layer_norm(x, (D_LLM,), W[hm + "layer_norm2.weight"], W[hm + "layer_norm2.bias"], 1e-5)
    return F.linear(x, W[hm + "fc2.weight"], W[hm + "fc2.bias"])


def unnormalize(normalized: np.ndarray, hi, lo, mask=None) -> np.ndarray:
    """_unnormalize_actions (MP:786-805) in float64 numpy; (hi, lo) = (q99, q01) or (max, min)."""
    hi, lo = np.array(hi), np.array(lo)
    if mask is None:
        mask = np.ones_like(lo, dtype=bool)
    return np.where(mask, 0.5 * (normalized + 1) * (hi - lo + 1e-8) + lo, normalized)


# =====================================================================================================
# the whole path
# =====================================================================================================
@torch.no_grad()
def predict_action_batch(W, cfg: OracleConfig, pixel_values, input_ids, proprio, dtype=torch.float32,
                         keep_taps=False):
    """Batched OpenVLAForActionPrediction.predict_action (MP:892-972).  Returns a dict with `normalized`
    (B, T, A) fp32, `last_ha` (B, 64, D), the integer tensors, and (keep_taps) every stage boundary."""
    Wd = {k: v.to(dtype) for k, v in W.items()}
    ext, labels, mask = prepare_inputs(input_ids)
    pix = pixel_values.to(dtype)
    patches = vision_backbone(pix, Wd, cfg)
    projected = projector(patches, Wd)
    x = assemble(ext, mask, projected, Wd)
    states = qwen_prefill(x, Wd, cfg)
    NP, L = cfg.num_patches, input_ids.shape[1]
    multi = gather_hidden(states, NP, L)
    head_dtype = torch.bfloat16 if dtype == torch.bfloat16 else dtype
    taps = {} if keep_taps else None
    act = policy_head(multi.to(head_dtype), torch.as_tensor(proprio).to(dtype), Wd, cfg, NP, head_dtype, taps)
    out = {
        "ext_ids": ext, "labels": labels, "mask": mask, "aq_index": aq_index_from_mask(mask),
        "normalized": act.reshape(-1, cfg.chunk_len, cfg.action_dim).float(),
        "last_ha": states[-1][:, NP + L - 1: NP + L - 1 + NUM_TOKENS],
    }
    if keep_taps:
        out.update({"patches": patches, "projected": projected, "llm_in": x, "multi": multi})
        for i, s in enumerate(states):
            out[f"hidden.{i}"] = s
        out.update(taps)
    return out
