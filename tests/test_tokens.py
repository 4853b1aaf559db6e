"""Host-side integer logic (vla_adapter_b200/tokens.py) against the oracle's restatement of MP:748-784 /
TU:8-41, including edge cases: one-token prompts, long prompts, batches, malformed masks."""
import pytest
import torch

from oracle import vla_oracle as O
from vla_adapter_b200 import tokens


@pytest.mark.parametrize("B,L", [(1, 1), (1, 2), (3, 17), (2, 56), (64, 48), (1, 512)])
def test_build_matches_oracle(B, L):
    g = torch.Generator().manual_seed(B * 1000 + L)
    ids = torch.randint(3, 151643, (B, L), generator=g)
    ext, labels, mask, aq, ext_mask = tokens.build(ids, None, 7)
    oext, olabels, omask = O.prepare_inputs(ids)
    assert torch.equal(ext, oext) and torch.equal(labels, olabels) and torch.equal(mask, omask)
    assert torch.equal(aq, O.aq_index_from_mask(omask))
    assert ext.shape == (B, L + 65) and ext.dtype == torch.int64 and aq.dtype == torch.int32
    assert (ext[:, L:L + 64] == 1).all() and (ext[:, -1] == tokens.STOP_INDEX).all()
    assert mask[:, L:L + 64].all() and not mask[:, :L].any() and not mask[:, -1].any()
    assert torch.equal(aq[:, L:L + 64], torch.arange(64, dtype=torch.int32).expand(B, 64))
    assert (aq[:, :L] == -1).all() and (aq[:, -1] == -1).all()
    assert ext_mask.shape == ext.shape and bool(ext_mask.all())


@pytest.mark.parametrize("action_dim", [1, 7, 14, 64])
def test_mask_union_independent_of_action_dim(action_dim):
    ids = torch.randint(3, 1000, (2, 9))
    _, labels, mask, _, _ = tokens.build(ids, None, action_dim)
    cur = tokens.get_current_action_mask(labels, action_dim)
    nxt = tokens.get_next_actions_mask(labels, action_dim)
    assert not (cur & nxt).any() and torch.equal(cur | nxt, mask)
    assert int(cur.sum()) == 2 * min(action_dim, 64)


def test_action_query_index_rejects_ragged_masks():
    m = torch.zeros(2, 80, dtype=torch.bool)
    m[0, 10:74] = True
    m[1, 10:73] = True
    with pytest.raises(ValueError):
        tokens.action_query_index(m)


def test_build_rejects_bad_rank():
    with pytest.raises(ValueError):
        tokens.build(torch.zeros(5, dtype=torch.int64), None, 7)
