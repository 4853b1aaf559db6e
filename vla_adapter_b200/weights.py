"""Shape manifest of the tensors the engine reads, keyed by the reference state_dict names
(vla. / head. / proprio. prefixes), and a device-side random initialiser for benchmarks.

Shapes follow pretrained_models/configs/config.json (text_config), timm's
vit_large_patch14_reg4_dinov2 / vit_so400m_patch14_siglip_224 (configuration_prismatic.py:36), the fused
projector (modeling_prismatic.py:254-257), MLPResNet (action_heads.py:84-108) and ProprioProjector
(projectors.py:15-16)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

D_DINO, F_DINO = 1024, 4096
D_SIG, F_SIG = 1152, 4304
D_LLM, I_LLM, HQ, HKV, HD = 896, 4864, 14, 2, 64


def weight_shapes(n_images=2, action_dim=7, proprio_dim=8, pro=False, dino_depth=24, siglip_depth=27,
                  llm_layers=24, vocab_size=151936) -> Dict[str, Tuple[int, ...]]:
    S: Dict[str, Tuple[int, ...]] = {}

    def lin(name, out_f, in_f, bias=True):
        S[name + ".weight"] = (out_f, in_f)
        if bias:
            S[name + ".bias"] = (out_f,)

    def ln(name, dim, bias=True):
        S[name + ".weight"] = (dim,)
        if bias:
            S[name + ".bias"] = (dim,)

    def tower(pfx, D, Fh, depth, dino):
        S[pfx + "patch_embed.proj.weight"] = (D, 3, 14, 14)
        S[pfx + "patch_embed.proj.bias"] = (D,)
        S[pfx + "pos_embed"] = (1, 256, D)
        if dino:
            S[pfx + "cls_token"] = (1, 1, D)
            S[pfx + "reg_token"] = (1, 4, D)
        for i in range(depth - 1):  # the last block never reaches the output (modeling_prismatic.py:141-142)
            b = f"{pfx}blocks.{i}."
            ln(b + "norm1", D)
            lin(b + "attn.qkv", 3 * D, D)
            lin(b + "attn.proj", D, D)
            ln(b + "norm2", D)
            lin(b + "mlp.fc1", Fh, D)
            lin(b + "mlp.fc2", D, Fh)
            if dino:
                S[b + "ls1.scale_factor"] = (D,)
                S[b + "ls2.scale_factor"] = (D,)

    tower("vla.vision_backbone.featurizer.", D_DINO, F_DINO, dino_depth, True)
    tower("vla.vision_backbone.fused_featurizer.", D_SIG, F_SIG, siglip_depth, False)
    lin("vla.projector.fc1", 4 * (D_DINO + D_SIG), D_DINO + D_SIG)
    lin("vla.projector.fc2", D_LLM, 4 * (D_DINO + D_SIG))
    lin("vla.projector.fc3", D_LLM, D_LLM)
    lm = "vla.language_model.model."
    S[lm + "embed_tokens.weight"] = (vocab_size, D_LLM)
    S["vla.action_queries.weight"] = (64, D_LLM)
    for i in range(llm_layers):
        b = f"{lm}layers.{i}."
        ln(b + "input_layernorm", D_LLM, bias=False)
        lin(b + "self_attn.q_proj", HQ * HD, D_LLM)
        lin(b + "self_attn.k_proj", HKV * HD, D_LLM)
        lin(b + "self_attn.v_proj", HKV * HD, D_LLM)
        lin(b + "self_attn.o_proj", D_LLM, D_LLM, bias=False)
        ln(b + "post_attention_layernorm", D_LLM, bias=False)
        lin(b + "mlp.gate_proj", I_LLM, D_LLM, bias=False)
        lin(b + "mlp.up_proj", I_LLM, D_LLM, bias=False)
        lin(b + "mlp.down_proj", D_LLM, I_LLM, bias=False)
    ln(lm + "norm", D_LLM, bias=False)
    hm = "head.model."
    ln(hm + "layer_norm1", action_dim * D_LLM)
    lin(hm + "fc1", D_LLM, action_dim * D_LLM)
    names = ["q_proj", "k_self", "v_self", "k_adapter", "v_adapter", "k_task", "v_task", "o_proj"] if pro else \
            ["q_proj", "k_proj", "v_proj", "o_proj"]
    for i in range(24):
        b = f"{hm}mlp_resnet_blocks.{i}."
        ln(b + "ffn.0", D_LLM)
        lin(b + "ffn.1", D_LLM, D_LLM)
        for n in names:
            lin(b + n, D_LLM, D_LLM)
        S[b + "gating_factor"] = (1,)
    ln(hm + "layer_norm2", D_LLM)
    lin(hm + "fc2", action_dim, D_LLM)
    lin("proprio.fc1", D_LLM, proprio_dim)
    lin("proprio.fc2", D_LLM, D_LLM)
    return S


def load_random_weights(engine, seed: int = 0, **shape_kw) -> int:
    """Random-init weights of the named architecture, generated ON THE DEVICE tensor by tensor and handed to
    the engine (benchmarks have no checkpoint to load: there is no network).  Returns the parameter count."""
    g = torch.Generator(device=engine.device).manual_seed(seed)
    total = 0
    for name, shape in weight_shapes(**shape_kw).items():
        numel = math.prod(shape)
        total += numel
        leaf = name.rsplit(".", 1)[-1]
        if len(shape) >= 2 and leaf == "weight":
            fan_in = numel // shape[0]
            t = torch.randn(shape, device=engine.device, generator=g) * (1.0 / math.sqrt(fan_in))
        elif leaf == "weight":  # norm scales
            t = 1.0 + 0.1 * torch.randn(shape, device=engine.device, generator=g)
        elif leaf == "scale_factor":
            t = 0.2 + 0.3 * torch.rand(shape, device=engine.device, generator=g)
        elif leaf in ("pos_embed", "cls_token", "reg_token"):
            t = 0.3 * torch.randn(shape, device=engine.device, generator=g)
        elif leaf == "gating_factor":
            t = torch.randn(shape, device=engine.device, generator=g)
        else:  # biases
            t = 0.05 * torch.randn(shape, device=engine.device, generator=g)
        engine.load_tensor(name, t.to(torch.bfloat16))
    return total
