"""CPU tests of the PyG-free training batch assembly (phoregen_b200/collate.py) against the unmodified reference transform
(datasets/transform.py:483-501) and against the dst-major edge order of the kernels."""
import os

import pytest
import torch

from oracle import phoregen_oracle as O
from phoregen_b200 import collate

HAVE_REF = os.path.isdir("/root/reference/models")


def _molecule(seed, n):
    g = torch.Generator().manual_seed(seed)
    pairs = torch.combinations(torch.arange(n), 2)
    pick = pairs[torch.rand(len(pairs), generator=g) < 0.3]
    attr = torch.randint(1, 5, (len(pick),), generator=g)
    edge_index = torch.cat([pick.T, pick.T.flip(0)], 1)                      # both directions, as the data set stores bonds
    return dict(x=torch.randint(0, 11, (n,), generator=g), pos=torch.randn(n, 3, generator=g), edge_index=edge_index,
                edge_attr=torch.cat([attr, attr]))


def _phore(seed, p):
    g = torch.Generator().manual_seed(seed)
    return dict(x=torch.rand(p, 18, generator=g), pos=torch.randn(p, 3, generator=g), norm=torch.randn(p, 3, generator=g))


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
@pytest.mark.parametrize("n", [2, 7, 19])
def test_featurize_matches_the_unmodified_reference(n):
    from oracle.shims.install import HeteroData, install
    install()
    from datasets.transform import FeaturizeLigandBond
    m = _molecule(n, n)
    d = HeteroData()
    d["ligand"].pos = m["pos"]
    d["ligand", "ligand"].edge_index = m["edge_index"]
    d["ligand", "ligand"].edge_attr = m["edge_attr"]
    want = FeaturizeLigandBond()(d)
    f_idx, f_attr = collate.featurize_ligand_bonds(n, m["edge_index"], m["edge_attr"])
    assert torch.equal(f_idx, want["ligand", "ligand"].f_edge_index) and torch.equal(f_attr, want["ligand", "ligand"].f_edge_attr)


def test_collated_batch_layout():
    sizes = [5, 2, 9]
    mols = [_molecule(i, n) for i, n in enumerate(sizes)]
    phs = [_phore(10 + i, 4 + i) for i in range(3)]
    b = collate.collate_training(mols, phs)
    lig, ll, ph = b["ligand"], b["ligand", "ligand"], b["phore"]
    assert b.num_graphs == 3 and lig.ptr.tolist() == [0, 5, 7, 16] and lig.batch.tolist() == sum(([g] * n for g, n in enumerate(sizes)), [])
    # the concatenated complete graph is the kernels' dst-major order (pg_plan_create edge_order = 1)
    ei, eb = O.full_edges_dst_major(torch.tensor(sizes))
    assert torch.equal(ll.f_edge_index, ei) and torch.equal(ll.f_edge_attr_batch, eb)
    assert ll.f_edge_attr.shape == (5 * 4 + 2 * 1 + 9 * 8,)
    # bond classes survive: entry (src, dst) of molecule 2
    m = mols[2]
    s, d_, c = int(m["edge_index"][0, 0]), int(m["edge_index"][1, 0]), int(m["edge_attr"][0])
    e = (ll.f_edge_index[0] == s + 7) & (ll.f_edge_index[1] == d_ + 7)
    assert int(e.sum()) == 1 and int(ll.f_edge_attr[e]) == c
    assert ph.batch.tolist() == [0] * 4 + [1] * 5 + [2] * 6 and ph.x.shape == (15, 18)
