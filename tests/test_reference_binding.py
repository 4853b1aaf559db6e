"""The documented drop-in binding: `VLAEngine.from_reference_modules(vla, action_head,
proprio_projector)` driven by LIVE modules of the unmodified reference (instantiated through oracle/ref_shim.py).
On the CPU box the engine below the Python layer is replaced by a recorder: what is checked is the read-out of every
shape parameter from the modules, the tensor names / shapes / dtypes handed to vla_load_tensor, and that the set
covers everything the engine's own weight list (oracle.make_weights) expects.  With a GPU the same call builds a real
engine and its actions are compared with the reference's own predict_action on the same modules."""
import numpy as np
import pytest
import torch

from oracle import ref_shim as R
from oracle import vla_oracle as O
from vla_adapter_b200.engine import VLAEngine

pytestmark = pytest.mark.skipif(not R.available(), reason="the reference sources are only mounted in the build container")

STATS = {"synthetic": {"action": {"q01": [-1.0] * 7, "q99": [1.0] * 7, "mask": [True] * 6 + [False]}}}


class _Recorder(VLAEngine):
    def __init__(self, **kw):  # no library, no device
        self.kw, self.loaded, self.finalized = kw, {}, False

    def load_tensor(self, name, t):
        assert name not in self.loaded, f"{name} handed over twice"
        self.loaded[name] = (tuple(t.shape), t.dtype)

    def finalize(self):
        self.finalized = True

    def close(self):
        pass


@pytest.mark.parametrize("pro,n_images", [(False, 2), (True, 2), (False, 1)])
def test_from_reference_modules_reads_out_live_modules(pro, n_images):
    cfg = O.OracleConfig(n_images=n_images, dino_depth=3, siglip_depth=2, vocab_size=1024, pro=pro)
    W = O.make_weights(cfg, seed=3)
    ns, vla, head, pp = R.build_reference(cfg, W, torch.bfloat16, norm_stats=STATS)
    eng = _Recorder.from_reference_modules(vla, head, pp, max_batch=4, max_prompt_len=40)
    assert eng.finalized
    want = dict(n_images=n_images, action_dim=7, chunk_len=8, proprio_dim=8, pro=pro, dino_depth=3, siglip_depth=2,
                llm_layers=24, vocab_size=1024, max_batch=4, max_prompt_len=40)
    for k, v in want.items():
        assert eng.kw[k] == v, (k, eng.kw[k], v)
    assert eng.kw["norm_stats"] == STATS          # the model's own statistics (MP:738)
    # every tensor the engine needs arrives under the engine's name with the reference's shape
    for k, v in W.items():
        assert k in eng.loaded, f"engine weight {k} was not handed over"
        assert eng.loaded[k][0] == tuple(v.shape), (k, eng.loaded[k][0], tuple(v.shape))
    assert all(n.split(".")[0] in ("vla", "head", "proprio") for n in eng.loaded)
    assert all(dt == torch.bfloat16 for _, dt in eng.loaded.values())   # as deployed: everything cast to bf16
    # nothing the path reads is missing, and what it does not read (lm_head, attention pool, film_gen) is only extra
    extra = sorted(set(eng.loaded) - set(W))
    assert all(any(p in n for p in ("lm_head", "attn_pool", "film_gen", "featurizer.norm.", "rope.", "blocks.2.",
                                    "blocks.1.")) for n in extra), extra[:10]


@pytest.mark.gpu
def test_from_reference_modules_engine_matches_reference_predict_action():
    cfg = O.OracleConfig(n_images=2, dino_depth=3, siglip_depth=3, vocab_size=1024, pro=True)
    W = O.make_weights(cfg, seed=5)
    ns, vla, head, pp = R.build_reference(cfg, W, torch.bfloat16, norm_stats=STATS)
    pix, ids, prop = O.make_inputs(cfg, 2, 22, seed=5)
    acts, hids = R.reference_predict_action(ns, vla, head, pp, pix, ids, prop, "synthetic")
    eng = VLAEngine.from_reference_modules(vla, head, pp, max_batch=2, max_prompt_len=22)
    for b in range(2):
        a, h = eng.predict_action(ids[b:b + 1], "synthetic", prop[b].numpy(), proprio_projector=pp, action_head=head,
                                  pixel_values=pix[b:b + 1], attention_mask=torch.ones_like(ids[b:b + 1]))
        assert a.shape == acts[b].shape and np.abs(a - acts[b]).max() <= 4e-2
        assert h.shape == hids[b].shape
    eng.close()
