"""Loads a VLA-Adapter checkpoint DIRECTORY into a VLAEngine without instantiating any reference module
(SURVEY.md 8f-2).  The directory layout is the reference's own (experiments/robot/openvla_utils.py):

    config.json                                  HF PrismaticConfig (text_config, image sizes, ...)          (:272-327)
    model.safetensors | model-0000i-of-0000n.safetensors (+ model.safetensors.index.json) | pytorch_model.bin
    action_head--<step>_checkpoint.pt            L1RegressionActionHead state dict, maybe "module."-prefixed (:412-453)
    proprio_projector--<step>_checkpoint.pt      ProprioProjector state dict                                 (:482-539)
    dataset_statistics.json                      un-normalisation statistics                                 (:371-390)

Tensors are streamed one at a time (safetensors `safe_open`), renamed to the engine's `vla.` / `head.` / `proprio.`
prefixes and handed to `vla_load_tensor`; names the path never reads (lm_head, attention-pool head, FiLM generators,
the discarded last ViT block) are skipped by the engine itself.
"""
from __future__ import annotations

import itertools
import json
import os
from typing import Any, Callable, Dict, Iterator, Optional, Tuple

import torch


def find_checkpoint_file(ckpt_dir: str, pattern: str) -> str:
    """Mirror of openvla_utils.find_checkpoint_file (:201-227): exactly one file containing `pattern` and
    "checkpoint"."""
    if not os.path.isdir(ckpt_dir):
        raise AssertionError(f"Checkpoint path must be a directory: {ckpt_dir}")
    hits = [os.path.join(ckpt_dir, f) for f in sorted(os.listdir(ckpt_dir)) if pattern in f and "checkpoint" in f]
    if len(hits) != 1:
        raise AssertionError(f"Expected exactly 1 {pattern} checkpoint but found {len(hits)} in directory: {ckpt_dir}")
    return hits[0]


def strip_module_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """openvla_utils.load_component_state_dict (:230-250): DDP-trained components carry a "module." prefix."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def iter_vla_tensors(ckpt_dir: str) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yields (name, tensor) of the HF VLM weights, whichever serialisation the directory holds."""
    index = os.path.join(ckpt_dir, "model.safetensors.index.json")
    single = os.path.join(ckpt_dir, "model.safetensors")
    if os.path.isfile(index) or os.path.isfile(single):
        from safetensors import safe_open

        if os.path.isfile(index):
            with open(index) as f:
                shards = sorted(set(json.load(f)["weight_map"].values()))
        else:
            shards = ["model.safetensors"]
        for shard in shards:
            with safe_open(os.path.join(ckpt_dir, shard), framework="pt", device="cpu") as f:
                for name in f.keys():
                    yield name, f.get_tensor(name)
        return
    bin_path = os.path.join(ckpt_dir, "pytorch_model.bin")
    if os.path.isfile(bin_path):
        for name, t in torch.load(bin_path, map_location="cpu", weights_only=True).items():
            yield name, t
        return
    raise FileNotFoundError(f"no model.safetensors[.index.json] or pytorch_model.bin in {ckpt_dir}")


def read_config(ckpt_dir: str) -> Dict[str, Any]:
    """Engine shape parameters from config.json (+ the head checkpoint for the variant / action dim)."""
    with open(os.path.join(ckpt_dir, "config.json")) as f:
        cfg = json.load(f)
    text = cfg.get("text_config", {})
    out = {"llm_layers": int(text.get("num_hidden_layers", 24)), "vocab_size": int(text.get("vocab_size", 151936))}
    if "pad_to_multiple_of" in cfg and "vocab_size" in text:
        m = int(cfg["pad_to_multiple_of"])
        out["vocab_size"] = (int(text["vocab_size"]) + m - 1) // m * m if m > 1 else int(text["vocab_size"])
    return out


def load_norm_stats(ckpt_dir: str) -> Optional[Dict[str, Any]]:
    """dataset_statistics.json when the directory has one (openvla_utils.py:371-390 overwrites vla.norm_stats with
    it), else the `norm_stats` entry of config.json, which is what the model itself carries
    (modeling_prismatic.py:738: `self.norm_stats = config.norm_stats`)."""
    p = os.path.join(ckpt_dir, "dataset_statistics.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    c = os.path.join(ckpt_dir, "config.json")
    if os.path.isfile(c):
        with open(c) as f:
            stats = json.load(f).get("norm_stats")
        if stats:
            return stats
    return None


def describe_head(head_sd: Dict[str, torch.Tensor]) -> Dict[str, Any]:
    """Variant and shape parameters the head checkpoint implies (action_heads.py:84-108, 286-330)."""
    pro = any(".k_self." in k for k in head_sd)
    action_dim = int(head_sd["model.fc2.weight"].shape[0])
    in_dim = int(head_sd["model.fc1.weight"].shape[1])
    hidden = int(head_sd["model.fc1.weight"].shape[0])
    if in_dim != action_dim * hidden:
        raise ValueError(f"head fc1 input {in_dim} != action_dim {action_dim} x hidden {hidden} (action_heads.py:37)")
    return {"pro": pro, "action_dim": action_dim}


def load_checkpoint(ckpt_dir: str, engine_factory: Callable[..., Any], n_images: int = 2, chunk_len: int = 8,
                    max_batch: int = 1, max_prompt_len: int = 64, vocab_size: Optional[int] = None,
                    dino_depth: int = 24, siglip_depth: int = 27, allow_missing_stats: bool = False, **engine_kw):
    """Builds and finalizes an engine from `ckpt_dir`.  `engine_factory` is VLAEngine (injected so that the host
    logic can be tested without a GPU).  The embedding table size is read from the checkpoint itself.  A directory
    without un-normalisation statistics is an error (the reference fails in _check_unnorm_key, MP:980-990) unless
    `allow_missing_stats` is set, in which case the engine returns NORMALISED actions."""
    head_sd = strip_module_prefix(torch.load(find_checkpoint_file(ckpt_dir, "action_head"), map_location="cpu",
                                             weights_only=True))
    prop_sd = strip_module_prefix(torch.load(find_checkpoint_file(ckpt_dir, "proprio_projector"), map_location="cpu",
                                             weights_only=True))
    hd = describe_head(head_sd)
    shape = read_config(ckpt_dir)
    proprio_dim = int(prop_sd["fc1.weight"].shape[1])
    vla_iter = iter_vla_tensors(ckpt_dir)
    pending = []
    if vocab_size is None:  # the embedding table knows best (configs pad the vocabulary, modeling_prismatic.py:381-384)
        for name, t in vla_iter:
            pending.append((name, t))
            if name.endswith("embed_tokens.weight"):
                vocab_size = int(t.shape[0])
                break
        if vocab_size is None:
            vocab_size = shape["vocab_size"]
    norm_stats = load_norm_stats(ckpt_dir)
    if norm_stats is None and not allow_missing_stats:
        raise FileNotFoundError(
            f"no dataset_statistics.json and no norm_stats in config.json under {ckpt_dir}: actions could not be "
            "un-normalised (pass allow_missing_stats=True to get normalised actions)")
    eng = engine_factory(n_images=n_images, chunk_len=chunk_len, action_dim=hd["action_dim"], proprio_dim=proprio_dim,
                         pro=hd["pro"], dino_depth=dino_depth, siglip_depth=siglip_depth, llm_layers=shape["llm_layers"],
                         vocab_size=vocab_size, max_batch=max_batch, max_prompt_len=max_prompt_len,
                         norm_stats=norm_stats, **engine_kw)
    n = 0
    for name, t in itertools.chain(pending, vla_iter):  # streamed: one tensor in host memory at a time
        if torch.is_tensor(t) and t.is_floating_point():
            eng.load_tensor("vla." + name, t)
            n += 1
    for prefix, sd in (("head.", head_sd), ("proprio.", prop_sd)):
        for k, v in sd.items():
            if torch.is_tensor(v) and v.is_floating_point():
                eng.load_tensor(prefix + k, v)
                n += 1
    eng.finalize()
    eng.loaded_tensors = n
    return eng
