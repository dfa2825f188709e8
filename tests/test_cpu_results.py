"""CPU tests of the per-molecule result views (phoregen_b200/results.py) against the reference's own
`unbatch_data` / `decode_data` (utils/sample_utils.py:57-132), imported unmodified where /root/reference is mounted,
and against a literal restatement of them otherwise."""
import os

import numpy as np
import pytest
import torch

from phoregen_b200 import results as R

HAVE_REF = os.path.isdir("/root/reference/models")


def _fake_results(seed, sizes, T=3, device="cpu"):
    """A `PhoreDiff.sample` result dict with random logits; a few atoms are forced to the mask class and a few edge
    logits to the mask / no-bond classes so that every branch of decode_data is taken."""
    g = torch.Generator().manual_seed(seed)
    n = torch.tensor(sizes)
    N, E = int(n.sum()), int((n * (n - 1)).sum())
    node = torch.randn(N, 12, generator=g)
    node[torch.randperm(N, generator=g)[: max(1, N // 6)], 11] = 9.0           # mask atoms
    edge = torch.randn(E, 6, generator=g)
    edge[torch.randperm(E, generator=g)[: E // 3], 0] = 9.0                     # no bond
    edge[torch.randperm(E, generator=g)[: E // 10], 5] = 9.5                    # mask bond
    pos = torch.randn(N, 3, generator=g)
    ei, eb, a0 = [], [], 0
    for i, k in enumerate(sizes):                                               # utils/sample_utils.py:40-54
        half = torch.triu_indices(k, k, offset=1)
        ei.append(torch.cat([half, half.flip(0)], 1) + a0)
        eb.append(torch.full((k * (k - 1),), i))
        a0 += k
    res = {"pred": [node, pos, edge],
           "traj": [torch.randn(T, N, 12, generator=g), torch.randn(T, N, 3, generator=g), torch.randn(T, E, 6, generator=g)],
           "lig_info": [n.int(), torch.repeat_interleave(torch.arange(len(sizes)), n), torch.cat(ei, 1), torch.cat(eb)]}
    return {k: [t.to(device) for t in v] for k, v in res.items()}


def _ref_unbatch(results, n_graphs, include_bond=True):          # literal restatement of utils/sample_utils.py:57-95
    pred, traj = results["pred"], results["traj"]
    batch_node, edge_index, batch_edge = results["lig_info"][1:4]
    out = []
    for i in range(n_graphs):
        ind_node, ind_edge = batch_node == i, batch_edge == i
        assert ind_node.sum() * (ind_node.sum() - 1) == ind_edge.sum()
        p = [pred[0][ind_node], pred[1][ind_node]] + ([pred[2][ind_edge]] if include_bond else [])
        t = [traj[0][:, ind_node], traj[1][:, ind_node]] + ([traj[2][:, ind_edge]] if include_bond else [])
        e = edge_index[:, ind_edge]
        out.append({"pred": p, "traj": t, "edge_index": e - ind_node.nonzero()[0].min()})
    return out


def _ref_decode(pred_info, edge_index, include_bond=True, num_bond_types=5):   # utils/sample_utils.py:98-132
    import torch.nn.functional as F
    atom_type = F.softmax(pred_info[0], dim=-1).argmax(dim=-1)
    ok = atom_type < len(R.ATOM_TYPES)
    if not ok.all():
        changer = -torch.ones(len(ok), dtype=torch.long)
        changer[ok] = torch.arange(ok.sum())
    element = [R.ATOM_TYPES[i] for i in atom_type[ok]]
    atom_pos = pred_info[1][ok]
    bond_type = bond_index = None
    if include_bond:
        edge_type = F.softmax(pred_info[2], dim=-1).argmax(dim=-1)
        is_bond = (edge_type > 0) & (edge_type < num_bond_types)
        bond_type, bond_index = edge_type[is_bond], edge_index[:, is_bond]
        if not ok.all():
            bond_index = changer[bond_index]
            bad = (bond_index < 0).any(dim=0)
            bond_index, bond_type = bond_index[:, ~bad], bond_type[~bad]
    return {"element": element, "atom_pos": atom_pos, "bond_type": bond_type, "bond_index": bond_index}


def _reference_functions():
    if not HAVE_REF:
        return _ref_unbatch, _ref_decode
    from oracle.shims.install import install
    install()
    from utils import sample_utils
    return sample_utils.unbatch_data, sample_utils.decode_data


def _same_decode(a, b):
    assert list(a["element"]) == list(b["element"])
    assert torch.equal(a["atom_pos"], b["atom_pos"])
    for k in ("bond_type", "bond_index"):
        if b[k] is None:
            assert a[k] is None
        else:
            assert a[k].shape == b[k].shape and torch.equal(a[k].long(), b[k].long()), k


@pytest.mark.parametrize("sizes", [[3, 7, 2, 12], [30] * 6, [2, 2, 2], [9]])
@pytest.mark.parametrize("include_bond", [True, False])
def test_unbatch_and_decode_match_the_reference(sizes, include_bond):
    ref_unbatch, ref_decode = _reference_functions()
    res = _fake_results(11 + len(sizes), sizes)
    G = len(sizes)
    want = ref_unbatch(res, G, include_bond=include_bond)
    got = R.unbatch_data(res, G, include_bond=include_bond)
    batch = R.decode_batch(res, G, include_bond=include_bond)
    assert len(got) == len(want) == len(batch) == G
    for g, w, bd in zip(got, want, batch):
        assert torch.equal(g["edge_index"], w["edge_index"])
        for key in ("pred", "traj"):
            assert len(g[key]) == len(w[key])
            for x, y in zip(g[key], w[key]):
                assert torch.equal(x, y)
        wd = ref_decode(w["pred"], w["edge_index"], include_bond=include_bond)
        _same_decode(R.decode_data(g["pred"], g["edge_index"], include_bond=include_bond), wd)
        _same_decode(bd, wd)


def test_restatement_used_without_the_reference_is_the_reference():
    """Pins the literal restatement above against the unmodified reference functions when they are importable."""
    if not HAVE_REF:
        pytest.skip("/root/reference not mounted")
    ref_unbatch, ref_decode = _reference_functions()
    res = _fake_results(5, [4, 6, 3])
    for a, b in zip(_ref_unbatch(res, 3), ref_unbatch(res, 3)):
        assert torch.equal(a["edge_index"], b["edge_index"])
        _same_decode(_ref_decode(a["pred"], a["edge_index"]), ref_decode(b["pred"], b["edge_index"]))


def test_layout_violations_are_rejected():
    res = _fake_results(3, [4, 5])
    with pytest.raises(ValueError):
        R.unbatch_data(res, 3)
    res["lig_info"][3] = res["lig_info"][3][:-2]
    with pytest.raises(AssertionError):
        R.unbatch_data(res, 2)
