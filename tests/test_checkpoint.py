"""Checkpoint-directory loader: host logic on a synthetic directory in the reference's layout (CPU),
and - on the GPU - the engine built from that directory against the same weights loaded directly."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vla_oracle as O
from vla_adapter_b200 import checkpoint as CK


def _write_ckpt(tmp, W, shards=2, ddp_prefix=True, stats=True):
    from safetensors.torch import save_file

    vla = {k[4:]: v.to(torch.bfloat16).contiguous() for k, v in W.items() if k.startswith("vla.")}
    vla["language_model.lm_head.weight"] = vla["language_model.model.embed_tokens.weight"].clone()   # tied, skipped
    names = sorted(vla)
    wm = {}
    for i in range(shards):
        part = {k: vla[k] for k in names[i::shards]}
        fn = f"model-{i + 1:05d}-of-{shards:05d}.safetensors"
        save_file(part, os.path.join(tmp, fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(os.path.join(tmp, "model.safetensors.index.json"), "w"))
    json.dump({"text_config": {"num_hidden_layers": 24, "vocab_size": 2048}, "pad_to_multiple_of": 64},
              open(os.path.join(tmp, "config.json"), "w"))
    pre = "module." if ddp_prefix else ""
    torch.save({pre + k[5:]: v for k, v in W.items() if k.startswith("head.")},
               os.path.join(tmp, "action_head--150000_checkpoint.pt"))
    torch.save({pre + k[8:]: v for k, v in W.items() if k.startswith("proprio.")},
               os.path.join(tmp, "proprio_projector--150000_checkpoint.pt"))
    if stats:
        json.dump({"synthetic": {"action": {"q01": [-1.0] * 7, "q99": [1.0] * 7, "mask": [True] * 6 + [False]}}},
                  open(os.path.join(tmp, "dataset_statistics.json"), "w"))


class _FakeEngine:
    """Records what the loader does (no GPU)."""

    def __init__(self, **kw):
        self.kw, self.loaded, self.finalized = kw, {}, False

    def load_tensor(self, name, t):
        assert not self.finalized
        self.loaded[name] = (tuple(t.shape), t.dtype)

    def finalize(self):
        self.finalized = True


@pytest.mark.parametrize("pro", [False, True])
def test_loader_host_logic(tmp_path, pro):
    cfg = O.OracleConfig(n_images=2, dino_depth=3, siglip_depth=3, vocab_size=2048, pro=pro)
    W = O.make_weights(cfg, seed=4)
    _write_ckpt(str(tmp_path), W)
    eng = CK.load_checkpoint(str(tmp_path), _FakeEngine, dino_depth=3, siglip_depth=3, max_batch=2, max_prompt_len=16)
    assert eng.finalized
    assert eng.kw["pro"] is pro and eng.kw["action_dim"] == 7 and eng.kw["proprio_dim"] == 8
    assert eng.kw["vocab_size"] == 2048 and eng.kw["llm_layers"] == 24
    assert eng.kw["norm_stats"]["synthetic"]["action"]["mask"][-1] is False
    # every engine weight arrives under the engine's names, DDP prefix stripped, nothing renamed wrongly
    for k, v in W.items():
        assert k in eng.loaded, k
        assert eng.loaded[k][0] == tuple(v.shape)
    assert "vla.language_model.lm_head.weight" in eng.loaded      # handed over; the engine ignores it by name
    assert not any(k.startswith("head.module.") or k.startswith("proprio.module.") for k in eng.loaded)


def test_loader_errors(tmp_path):
    with pytest.raises(AssertionError):
        CK.find_checkpoint_file(str(tmp_path / "missing"), "action_head")
    cfg = O.OracleConfig(n_images=1, dino_depth=2, siglip_depth=2, vocab_size=64, pro=False)
    W = O.make_weights(cfg, seed=0)
    _write_ckpt(str(tmp_path), W, shards=1, ddp_prefix=False, stats=False)
    open(os.path.join(str(tmp_path), "action_head--2_checkpoint.pt"), "wb").close()      # a second match
    with pytest.raises(AssertionError):
        CK.load_checkpoint(str(tmp_path), _FakeEngine)
    os.remove(os.path.join(str(tmp_path), "action_head--2_checkpoint.pt"))
    # no dataset_statistics.json and no norm_stats in config.json: the reference would fail in _check_unnorm_key
    with pytest.raises(FileNotFoundError, match="un-normalised"):
        CK.load_checkpoint(str(tmp_path), _FakeEngine, n_images=1, dino_depth=2, siglip_depth=2)
    eng = CK.load_checkpoint(str(tmp_path), _FakeEngine, n_images=1, dino_depth=2, siglip_depth=2,
                             allow_missing_stats=True)
    assert eng.kw["norm_stats"] is None and eng.kw["vocab_size"] == 64
    # ... but config.json's own norm_stats (what the model carries, MP:738) are the fallback
    cfg_path = os.path.join(str(tmp_path), "config.json")
    c = json.load(open(cfg_path))
    c["norm_stats"] = {"from_config": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
    json.dump(c, open(cfg_path, "w"))
    eng = CK.load_checkpoint(str(tmp_path), _FakeEngine, n_images=1, dino_depth=2, siglip_depth=2)
    assert list(eng.kw["norm_stats"]) == ["from_config"]


@pytest.mark.gpu
def test_engine_from_checkpoint_dir_matches_direct_load(tmp_path):
    from vla_adapter_b200.engine import VLAEngine

    cfg = O.OracleConfig(n_images=2, dino_depth=3, siglip_depth=3, vocab_size=2048, pro=True)
    W = O.make_weights(cfg, seed=6)
    # the checkpoint stores the VLM in bf16 (like the published ones): compare against the same rounding
    Wb = {k: (v.to(torch.bfloat16).float() if k.startswith("vla.") else v) for k, v in W.items()}
    _write_ckpt(str(tmp_path), W)
    pix, ids, prop = O.make_inputs(cfg, 2, 18, seed=6)
    a = CK.load_checkpoint(str(tmp_path), VLAEngine, dino_depth=3, siglip_depth=3, max_batch=2, max_prompt_len=18)
    act_a, norm_a = a.predict_action_batch(ids, None, pix, prop, unnorm_key="synthetic")
    a.close()
    b = VLAEngine(n_images=2, pro=True, dino_depth=3, siglip_depth=3, vocab_size=2048, max_batch=2, max_prompt_len=18,
                  norm_stats=json.load(open(os.path.join(str(tmp_path), "dataset_statistics.json"))))
    b.load_flat(Wb)
    b.finalize()
    act_b, norm_b = b.predict_action_batch(ids, None, pix, prop, unnorm_key="synthetic")
    b.close()
    assert np.array_equal(norm_a, norm_b) and np.array_equal(act_a, act_b)
