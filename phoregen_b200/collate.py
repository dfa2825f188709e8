"""Training-side batch assembly without PyG (host side): `FeaturizeLigandBond`
(datasets/transform.py:483-501) and the collation that `DataLoader(..., follow_batch=['f_edge_attr'])` performs
(run/run.py:96-103), producing the batch `PhoreDiff.compute_loss` reads (phoregen_b200/testing.py:TrainBatch).

The complete-graph edge list is dst-major with src ascending and the self pair removed - the order the kernels use
internally (`edge_order=1` of `pg_plan_create`), so no permutation is needed at run time.
"""
import torch

from .testing import TrainBatch


def featurize_ligand_bonds(n_atoms, edge_index, edge_attr):
    """datasets/transform.py:488-501 -> (f_edge_index [2, n(n-1)], f_edge_attr [n(n-1)]): every ordered pair (src, dst),
    dst-major; class 0 = no bond, otherwise the class of the (src, dst) entry of the bond list."""
    dst = torch.repeat_interleave(torch.arange(n_atoms), n_atoms)
    src = torch.arange(n_atoms).repeat(n_atoms)
    keep = dst != src
    f_edge_index = torch.stack([src[keep], dst[keep]], dim=0)
    bond = torch.zeros(n_atoms, n_atoms, dtype=torch.long)
    bond[edge_index[0], edge_index[1]] = edge_attr
    return f_edge_index, bond[f_edge_index[0], f_edge_index[1]]


def collate_training(molecules, phores):
    """molecules: list of dict(x [n] atom classes, pos [n,3], edge_index [2,b], edge_attr [b]);
    phores: list of dict(x [p,18], pos [p,3], norm [p,3]) (e.g. `phore_io.parse_phore_file(...)['phore']`), one per molecule.
    -> TrainBatch with node / edge / pharmacophore tensors concatenated, indices offset, and the `batch` / `ptr` vectors."""
    assert len(molecules) == len(phores) and len(molecules) > 0
    xs, ps, bn, fe, fa, fb, ei, ptr = [], [], [], [], [], [], [], [0]
    px, pp, pn, pb = [], [], [], []
    a0 = 0
    for g, (m, ph) in enumerate(zip(molecules, phores)):
        n = m["pos"].shape[0]
        f_idx, f_attr = featurize_ligand_bonds(n, m["edge_index"], m["edge_attr"])
        xs.append(m["x"]); ps.append(m["pos"]); bn.append(torch.full((n,), g, dtype=torch.long))
        fe.append(f_idx + a0); fa.append(f_attr); fb.append(torch.full((f_attr.numel(),), g, dtype=torch.long))
        ei.append(m["edge_index"] + a0)
        a0 += n
        ptr.append(a0)
        px.append(ph["x"]); pp.append(ph["pos"]); pn.append(ph["norm"]); pb.append(torch.full((ph["x"].shape[0],), g, dtype=torch.long))
    return TrainBatch(
        ligand=dict(x=torch.cat(xs), pos=torch.cat(ps), batch=torch.cat(bn), ptr=torch.tensor(ptr)),
        bonds=dict(f_edge_attr=torch.cat(fa), f_edge_index=torch.cat(fe, 1), f_edge_attr_batch=torch.cat(fb), edge_index=torch.cat(ei, 1)),
        phore=dict(x=torch.cat(px), pos=torch.cat(pp), norm=torch.cat(pn), batch=torch.cat(pb)),
        num_graphs=len(molecules))
