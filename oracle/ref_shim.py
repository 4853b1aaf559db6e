"""Imports the UNMODIFIED reference modules from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY (oracle/ rules apply).  Used by oracle/make_golden.py to produce the golden vectors
under tests/golden/ and by tests/test_oracle_vs_reference.py (skipped when /root/reference is absent, i.e. on
the GPU box).  No reference source is copied: the files are executed where they lie.

The reference cannot be imported as a package here (prismatic/__init__.py pulls draccus, timm, tensorflow,
all missing), so
  * empty namespace stubs are registered for the prismatic.* packages and the leaf files
        prismatic/vla/constants.py, prismatic/training/train_utils.py, prismatic/models/action_heads.py,
        prismatic/models/projectors.py, prismatic/extern/hf/{configuration,modeling}_prismatic.py
    are loaded under their real module names, unmodified;
  * `timm` (0.9.10, un-vendored third-party) is replaced by a stand-in that builds a VisionTransformer with
    timm's parameter names and timm's get_intermediate_layers semantics (restated in the reference itself at
    prismatic/models/film_vit_wrapper.py:114-168); the ViT arithmetic is therefore pinned separately against
    transformers' independent Dinov2WithRegisters / Siglip implementations (tests/test_oracle_vit_hf.py);
  * two transformers-5.5 incompatibilities of the 4.40.1-era file are neutralised without touching its
    logic: `tie_weights` gets **kwargs, and `_supports_sdpa` is not evaluated before language_model exists
    (config._attn_implementation is set explicitly).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("VLA_REFERENCE_ROOT", "/root/reference")

# depth overrides for reduced-depth parity cases: {"dino": int, "siglip": int}
VIT_DEPTH = {"dino": 24, "siglip": 27}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "prismatic/extern/hf/modeling_prismatic.py"))


# ------------------------------------------------------------------------------------------------------
# timm stand-in (names and call semantics of timm 0.9.10's VisionTransformer; see module docstring)
# ------------------------------------------------------------------------------------------------------
class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x.mul_(self.gamma) if self.inplace else x * self.gamma


class _Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, num_heads, hidden, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads)
        self.ls1 = LayerScale(dim, init_values) if init_values else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)
        self.ls2 = LayerScale(dim, init_values) if init_values else nn.Identity()

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, dim):
        super().__init__()
        self.grid_size = (img_size // patch, img_size // patch)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """timm VisionTransformer subset: no_embed_class=True position embedding (cls/reg tokens are
    concatenated AFTER the pos-add), pre-norm blocks, get_intermediate_layers."""

    def __init__(self, img_size, dim, depth, heads, hidden, init_values, class_token, reg_tokens):
        super().__init__()
        self.embed_dim = dim
        self.patch_embed = _PatchEmbed(img_size, 14, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim)) if class_token else None
        self.reg_token = nn.Parameter(torch.zeros(1, reg_tokens, dim)) if reg_tokens else None
        self.num_prefix_tokens = (1 if class_token else 0) + reg_tokens
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches, dim) * 0.02)
        self.blocks = nn.Sequential(*[_Block(dim, heads, hidden, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def _pos_embed(self, x):
        x = x + self.pos_embed
        pre = []
        if self.cls_token is not None:
            pre.append(self.cls_token.expand(x.shape[0], -1, -1))
        if self.reg_token is not None:
            pre.append(self.reg_token.expand(x.shape[0], -1, -1))
        return torch.cat(pre + [x], dim=1) if pre else x

    def _intermediate_layers(self, x, n=1):
        outputs, num_blocks = [], len(self.blocks)
        take = set(range(num_blocks - n, num_blocks) if isinstance(n, int) else n)
        x = self._pos_embed(self.patch_embed(x))
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i in take:
                outputs.append(x)
        return outputs

    def get_intermediate_layers(self, x, n=1, reshape=False, return_prefix_tokens=False, norm=False):
        outputs = self._intermediate_layers(x, n)
        if norm:
            outputs = [self.norm(o) for o in outputs]
        outputs = [o[:, self.num_prefix_tokens:] for o in outputs]
        return tuple(outputs)


def _create_model(model_id, pretrained=False, num_classes=0, img_size=224, act_layer=None):
    assert not pretrained and act_layer is None
    if model_id == "vit_large_patch14_reg4_dinov2.lvd142m":
        return VisionTransformer(img_size, 1024, VIT_DEPTH["dino"], 16, 4096, 1e-5, True, 4)
    if model_id == "vit_so400m_patch14_siglip_224":
        return VisionTransformer(img_size, 1152, VIT_DEPTH["siglip"], 16, 4304, None, False, 0)
    raise ValueError(f"timm stand-in does not know {model_id}")


def _install_timm_standin():
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_vla_standin", False):
        return  # a real timm is importable: use it
    timm = types.ModuleType("timm")
    timm.__version__ = "0.9.10"
    timm._vla_standin = True
    timm.create_model = _create_model
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.LayerScale = LayerScale
    vt.VisionTransformer = VisionTransformer
    models.vision_transformer = vt
    timm.models = models
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vt})


# ------------------------------------------------------------------------------------------------------
# loading the reference leaf modules under their own names
# ------------------------------------------------------------------------------------------------------
_LOADED = {}


def _load(modname: str, relpath: str):
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load(platform: str = "libero"):
    """Returns a namespace with the reference modules K, TU, AH, PJ, CP, MP (loaded once per process; the
    platform constants are baked at first import, like in the reference: constants.py:58-91)."""
    if _LOADED:
        if _LOADED["platform"] != platform:
            raise RuntimeError("reference constants are import-time globals; use a fresh process per platform")
        return _LOADED["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import transformers  # noqa: F401  (must precede MP)

    _install_timm_standin()
    for pkg in ("prismatic", "prismatic.models", "prismatic.vla", "prismatic.training", "prismatic.extern",
                "prismatic.extern.hf"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []  # namespace stub: no __init__ is executed
            sys.modules[pkg] = m
    argv = list(sys.argv)
    sys.argv = [argv[0] if argv else "x", platform]
    try:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):  # constants.py prints the platform table
            K = _load("prismatic.vla.constants", "prismatic/vla/constants.py")
    finally:
        sys.argv = argv
    TU = _load("prismatic.training.train_utils", "prismatic/training/train_utils.py")
    AH = _load("prismatic.models.action_heads", "prismatic/models/action_heads.py")
    PJ = _load("prismatic.models.projectors", "prismatic/models/projectors.py")
    CP = _load("prismatic.extern.hf.configuration_prismatic", "prismatic/extern/hf/configuration_prismatic.py")
    MP = _load("prismatic.extern.hf.modeling_prismatic", "prismatic/extern/hf/modeling_prismatic.py")
    ns = types.SimpleNamespace(K=K, TU=TU, AH=AH, PJ=PJ, CP=CP, MP=MP)
    _LOADED.update(platform=platform, ns=ns)
    return ns


def text_config_dict(vocab_size: int, llm_layers: int = 24) -> dict:
    """pretrained_models/configs/config.json text_config (Qwen2.5-0.5B), vocabulary/depth overridable."""
    return dict(hidden_size=896, intermediate_size=4864, num_hidden_layers=llm_layers, num_attention_heads=14,
                num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1000000.0, vocab_size=vocab_size,
                max_position_embeddings=32768, tie_word_embeddings=True, hidden_act="silu",
                use_sliding_window=False, attention_dropout=0.0)


def build_reference(cfg, W: dict, dtype=torch.bfloat16, norm_stats=None):
    """Instantiates OpenVLAForActionPrediction + L1RegressionActionHead + ProprioProjector from the reference
    classes and loads the oracle's seeded weights `W` (keys vla./head./proprio.) into them."""
    ns = load("libero")
    K = ns.K
    assert (cfg.chunk_len, cfg.action_dim, cfg.proprio_dim) == (K.NUM_ACTIONS_CHUNK, K.ACTION_DIM, K.PROPRIO_DIM), \
        "reference constants are import-time globals (LIBERO: 8/7/8)"
    VIT_DEPTH["dino"], VIT_DEPTH["siglip"] = cfg.dino_depth, cfg.siglip_depth
    MP = ns.MP
    if not getattr(MP.PrismaticForConditionalGeneration, "_vla_tie_patched", False):
        orig_tie = MP.PrismaticForConditionalGeneration.tie_weights

        def tie_weights(self, *a, **kw):  # transformers 5.5 passes recompute_mapping=...
            return orig_tie(self)

        MP.PrismaticForConditionalGeneration.tie_weights = tie_weights
        MP.PrismaticForConditionalGeneration._vla_tie_patched = True
    conf = ns.CP.OpenVLAConfig(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="qwen25-0_5b-extra",
                               arch_specifier="no-align+fused-gelu-mlp", use_fused_vision_backbone=True,
                               image_resize_strategy="resize-naive",
                               text_config=text_config_dict(cfg.vocab_size, cfg.llm_layers), pad_token_id=0,
                               norm_stats=norm_stats)
    conf._attn_implementation = "eager"
    conf.text_config._attn_implementation = "sdpa"
    torch.manual_seed(0)
    vla = MP.OpenVLAForActionPrediction(conf)
    vla.vision_backbone.set_num_images_in_input(cfg.n_images)
    head = ns.AH.L1RegressionActionHead(input_dim=896, hidden_dim=896, action_dim=cfg.action_dim,
                                        num_task_tokens=cfg.num_patches, use_pro_version=cfg.pro)
    pp = ns.PJ.ProprioProjector(llm_dim=896, proprio_dim=cfg.proprio_dim)

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in W.items() if k.startswith(prefix)}

    sd = sub("vla.")
    sd["language_model.lm_head.weight"] = sd["language_model.model.embed_tokens.weight"]
    missing, unexpected = vla.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "attn_pool" not in m and "rotary_emb" not in m]
    assert not missing and not unexpected, (missing[:8], unexpected[:8])
    m2, u2 = head.load_state_dict(sub("head."), strict=False)
    m2 = [m for m in m2 if "film_gen" not in m and "rope" not in m]
    assert not m2 and not u2, (m2[:8], u2[:8])
    pp.load_state_dict(sub("proprio."))
    # The reference loads with from_pretrained(torch_dtype=bf16) (openvla_utils.py:288-300): weights become bf16
    # but Qwen2's rotary inv_freq buffer is created in fp32 and never cast.  A plain .to(bf16) would cast it, so
    # it is put back.  (The ACTION HEAD, in contrast, really is cast with .to(torch.bfloat16), openvla_utils.py:515.)
    rot = vla.language_model.model.rotary_emb
    inv_freq = rot.inv_freq.clone()
    vla = vla.to(dtype).eval()
    rot.inv_freq = inv_freq
    if hasattr(rot, "original_inv_freq"):
        rot.original_inv_freq = inv_freq
    head = head.to(torch.bfloat16 if dtype == torch.bfloat16 else dtype).eval()
    pp = pp.to(torch.bfloat16 if dtype == torch.bfloat16 else dtype).eval()
    return ns, vla, head, pp


@torch.no_grad()
def reference_predict_action(ns, vla, head, pp, pixel_values, input_ids, proprio, unnorm_key=None):
    """The reference's own bs=1 entry point, one call per sample (MP:892-972)."""
    outs, hids = [], []
    for b in range(input_ids.shape[0]):
        ids = input_ids[b:b + 1]
        a, h = vla.predict_action(input_ids=ids, unnorm_key=unnorm_key, proprio=proprio[b].numpy(),
                                  proprio_projector=pp, action_head=head,
                                  pixel_values=pixel_values[b:b + 1].to(next(vla.parameters()).dtype),
                                  attention_mask=torch.ones_like(ids))
        outs.append(a)
        hids.append(h)
    return outs, hids
