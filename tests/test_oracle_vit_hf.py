"""timm==0.9.10 (pyproject.toml:45 of the reference) is not installable here, so the oracle's ViT restatement
(vit_tower / vit_block, following film_vit_wrapper.py:69-75, 124-137, 153-168) is pinned against the
INDEPENDENT implementations of the same architectures in transformers: Dinov2WithRegistersModel and
SiglipVisionModel (different parameter names, same math).  fp32, seeded weights, reduced depth."""
import pytest
import torch

from oracle import vla_oracle as O

transformers = pytest.importorskip("transformers")


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _copy_block(sd, W, b, hf, D, names):
    qkv_w, qkv_b = W[b + "attn.qkv.weight"].float(), W[b + "attn.qkv.bias"].float()
    for j, n in enumerate(names["qkv"]):
        sd[hf + n + ".weight"] = qkv_w[j * D:(j + 1) * D]
        sd[hf + n + ".bias"] = qkv_b[j * D:(j + 1) * D]
    for src, dst in names["map"].items():
        sd[hf + dst] = W[b + src].float()


def test_dinov2_reg4_tower_matches_hf():
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    depth = 3
    cfg = O.OracleConfig(n_images=1, dino_depth=depth, siglip_depth=2, vocab_size=64)
    W = O.make_weights(cfg, seed=5)
    pfx = "vla.vision_backbone.featurizer."
    hc = Dinov2WithRegistersConfig(hidden_size=1024, num_hidden_layers=depth, num_attention_heads=16, mlp_ratio=4,
                                   image_size=224, patch_size=14, num_register_tokens=4, layer_norm_eps=1e-6,
                                   layerscale_value=1.0, hidden_act="gelu", qkv_bias=True, use_swiglu_ffn=False)
    m = Dinov2WithRegistersModel(hc).eval()
    sd = m.state_dict()
    sd["embeddings.patch_embeddings.projection.weight"] = W[pfx + "patch_embed.proj.weight"].float()
    sd["embeddings.patch_embeddings.projection.bias"] = W[pfx + "patch_embed.proj.bias"].float()
    # timm no_embed_class=True: pos_embed covers the 256 patches only, cls/reg get none -> zero cls position in HF
    sd["embeddings.position_embeddings"] = torch.cat([torch.zeros(1, 1, 1024), W[pfx + "pos_embed"].float()], dim=1)
    sd["embeddings.cls_token"] = W[pfx + "cls_token"].float()
    sd["embeddings.register_tokens"] = W[pfx + "reg_token"].float()
    names = {"qkv": ["attention.attention.query", "attention.attention.key", "attention.attention.value"],
             "map": {"norm1.weight": "norm1.weight", "norm1.bias": "norm1.bias", "norm2.weight": "norm2.weight",
                     "norm2.bias": "norm2.bias", "attn.proj.weight": "attention.output.dense.weight",
                     "attn.proj.bias": "attention.output.dense.bias", "mlp.fc1.weight": "mlp.fc1.weight",
                     "mlp.fc1.bias": "mlp.fc1.bias", "mlp.fc2.weight": "mlp.fc2.weight", "mlp.fc2.bias": "mlp.fc2.bias",
                     "ls1.scale_factor": "layer_scale1.lambda1", "ls2.scale_factor": "layer_scale2.lambda1"}}
    for i in range(depth):
        _copy_block(sd, W, f"{pfx}blocks.{i}.", f"encoder.layer.{i}.", 1024, names)
    m.load_state_dict(sd)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        hs = m(pixel_values=img, output_hidden_states=True).hidden_states
        Wf = {k: v.float() for k, v in W.items()}
        mine = O.vit_tower(img, Wf, pfx, 1024, depth, True)
    ref = hs[depth - 1][:, 5:]  # output of block index depth-2, prefix (cls + 4 reg) stripped, no final norm
    assert mine.shape == ref.shape == (2, 256, 1024)
    assert _rel(mine, ref) < 1e-5


def test_siglip_so400m_tower_matches_hf():
    from transformers import SiglipVisionConfig, SiglipVisionModel

    depth = 3
    cfg = O.OracleConfig(n_images=1, dino_depth=2, siglip_depth=depth, vocab_size=64)
    W = O.make_weights(cfg, seed=6)
    pfx = "vla.vision_backbone.fused_featurizer."
    hc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=depth, num_attention_heads=16,
                            image_size=224, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu")
    m = SiglipVisionModel(hc).eval()
    sd = m.state_dict()
    p = "vision_model."
    sd[p + "embeddings.patch_embedding.weight"] = W[pfx + "patch_embed.proj.weight"].float()
    sd[p + "embeddings.patch_embedding.bias"] = W[pfx + "patch_embed.proj.bias"].float()
    sd[p + "embeddings.position_embedding.weight"] = W[pfx + "pos_embed"].float()[0]
    names = {"qkv": ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"],
             "map": {"norm1.weight": "layer_norm1.weight", "norm1.bias": "layer_norm1.bias",
                     "norm2.weight": "layer_norm2.weight", "norm2.bias": "layer_norm2.bias",
                     "attn.proj.weight": "self_attn.out_proj.weight", "attn.proj.bias": "self_attn.out_proj.bias",
                     "mlp.fc1.weight": "mlp.fc1.weight", "mlp.fc1.bias": "mlp.fc1.bias",
                     "mlp.fc2.weight": "mlp.fc2.weight", "mlp.fc2.bias": "mlp.fc2.bias"}}
    for i in range(depth):
        _copy_block(sd, W, f"{pfx}blocks.{i}.", f"{p}encoder.layers.{i}.", 1152, names)
    m.load_state_dict(sd)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        hs = m(pixel_values=img, output_hidden_states=True).hidden_states
        Wf = {k: v.float() for k, v in W.items()}
        mine = O.vit_tower(img, Wf, pfx, 1152, depth, False)
    ref = hs[depth - 1]
    assert mine.shape == ref.shape == (2, 256, 1152)
    assert _rel(mine, ref) < 1e-5
