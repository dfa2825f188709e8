"""Per-molecule views of a sampled batch: drop-ins for `unbatch_data` / `decode_data` of the reference
(utils/sample_utils.py:57-132; consumer sample_all.py:104-175) plus a batched decode.

The reference selects every molecule with two boolean masks over the whole batch (`batch_node == i`, `batch_edge == i`):
O(G * (N + E)) work and 2 G device round trips.  Nodes and edges of a molecule are contiguous in the layouts that
`make_edge_data` / `PhoreDiff.sample` produce (utils/sample_utils.py:40-54), so offsets from `num_atoms` give the same
slices in O(N + E).  Outputs are element-for-element identical to the reference's (tests/test_cpu_results.py).

These are host-side callers of the hot path; nothing here runs inside the reverse loop.
"""
import torch

ATOM_TYPES = [5, 6, 7, 8, 9, 14, 15, 16, 17, 35, 53]   # utils/sample_utils.py:17 (the mask atom is class 11, the last one)


class ClassTrajectory:
    """Categorical trajectory kept as class indices: the compact stand-in for the reference's `[T+1, rows, K]` one-hot f32
    trajectory tensors (models/diffusion.py:418-426).  It supports exactly what the reference's consumers do with them -
    `.cpu()` (sample_all.py:104), `[:, mask]` per molecule (utils/sample_utils.py:74-76), `[t]` and `len()` per time step
    (sample_all.py:138-143) - and materialises one-hot f32 only for the `[t]` slice that is read.  At configs[1] the dense
    edge trajectory is 21 GB; this holds 0.9 GB."""

    def __init__(self, classes, num_classes):
        assert classes.dim() == 2 and classes.dtype == torch.uint8
        self.classes, self.num_classes = classes, int(num_classes)

    # ---- tensor-like surface
    @property
    def shape(self):
        return torch.Size((*self.classes.shape, self.num_classes))

    @property
    def device(self):
        return self.classes.device

    @property
    def is_cuda(self):
        return self.classes.is_cuda

    def __len__(self):
        return self.classes.shape[0]

    def cpu(self):
        return ClassTrajectory(self.classes.cpu(), self.num_classes)

    def to(self, *args, **kwargs):
        kwargs.pop("dtype", None)
        args = [a for a in args if not isinstance(a, torch.dtype)]
        return ClassTrajectory(self.classes.to(*args, **kwargs), self.num_classes)

    def dense(self):
        """The reference's tensor: one-hot f32 [T+1, rows, K]."""
        return torch.nn.functional.one_hot(self.classes.long(), self.num_classes).float()

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            if len(idx) > 2:
                return self.dense()[idx]
            sub = self.classes[idx[0]]                       # time index first, then the row selection on the last dim
            if len(idx) == 2:
                sub = sub[..., idx[1]]
        else:
            sub = self.classes[idx]
        if sub.dim() == 2:
            return ClassTrajectory(sub, self.num_classes)
        return torch.nn.functional.one_hot(sub.long(), self.num_classes).float()      # one time step (or one row): dense


def _offsets(num_atoms):
    n = torch.as_tensor(num_atoms).detach().to("cpu", torch.int64)
    node_off = torch.zeros(n.numel() + 1, dtype=torch.int64)
    node_off[1:] = torch.cumsum(n, 0)
    edge_off = torch.zeros(n.numel() + 1, dtype=torch.int64)
    edge_off[1:] = torch.cumsum(n * (n - 1), 0)
    return n, node_off.tolist(), edge_off.tolist()


def unbatch_data(results, n_graphs, include_bond=True):
    """utils/sample_utils.py:57-95.  `results` is the dict returned by `PhoreDiff.sample`:
    pred = [node logits [N,12], pos [N,3], edge logits [E,6]], traj = the same three with time as dim 0,
    lig_info = [num_atoms [G], batch_node [N], edge_index [2,E], batch_edge [E]].
    Returns one dict per molecule: pred, traj, edge_index (re-based to the molecule's first atom)."""
    pred, traj = results["pred"], results["traj"]
    num_atoms, batch_node, edge_index, batch_edge = results["lig_info"][:4]
    n, node_off, edge_off = _offsets(num_atoms)
    if n.numel() != n_graphs:
        raise ValueError(f"lig_info holds {n.numel()} molecules, n_graphs = {n_graphs}")
    if node_off[-1] != batch_node.shape[0] or edge_off[-1] != batch_edge.shape[0]:
        # the reference asserts n*(n-1) == number of edges per molecule (utils/sample_utils.py:68)
        raise AssertionError("ligand layout is not the complete-graph layout of make_edge_data")
    out = []
    for i in range(n_graphs):
        a0, a1, e0, e1 = node_off[i], node_off[i + 1], edge_off[i], edge_off[i + 1]
        if include_bond:
            p = [pred[0][a0:a1], pred[1][a0:a1], pred[2][e0:e1]]
            t = [traj[0][:, a0:a1], traj[1][:, a0:a1], traj[2][:, e0:e1]]
        else:
            p = [pred[0][a0:a1], pred[1][a0:a1]]
            t = [traj[0][:, a0:a1], traj[1][:, a0:a1]]
        out.append({"pred": p, "traj": t, "edge_index": edge_index[:, e0:e1] - a0})
    return out


def decode_data(pred_info, edge_index, include_bond=True, num_bond_types=5):
    """utils/sample_utils.py:98-132: arg-max classes, mask atoms (class >= 11) dropped and bond indices re-numbered,
    bonds = edge classes 1..num_bond_types-1 between kept atoms."""
    atom_type = pred_info[0].argmax(dim=-1)               # softmax is monotone: same arg-max as the reference
    keep = atom_type < len(ATOM_TYPES)
    element = [ATOM_TYPES[i] for i in atom_type[keep].tolist()]
    atom_pos = pred_info[1][keep]
    bond_type = bond_index = None
    if include_bond:
        edge_type = pred_info[2].argmax(dim=-1)
        is_bond = (edge_type > 0) & (edge_type < num_bond_types)
        bond_type, bond_index = edge_type[is_bond], edge_index[:, is_bond]
        if not bool(keep.all()):
            renum = torch.full((keep.numel(),), -1, dtype=torch.long, device=keep.device)
            renum[keep] = torch.arange(int(keep.sum()), device=keep.device)
            bond_index = renum[bond_index]
            ok = ~(bond_index < 0).any(dim=0)
            bond_index, bond_type = bond_index[:, ok], bond_type[ok]
    return {"element": element, "atom_pos": atom_pos, "bond_type": bond_type, "bond_index": bond_index}


def decode_batch(results, n_graphs, include_bond=True, num_bond_types=5):
    """decode_data(unbatch_data(results)[i]['pred'], ...) for every molecule, with the arg-max, the mask filters and the
    re-numbering done once for the whole batch (one device pass, one device->host copy of uint8 classes and fp32
    positions) instead of once per molecule.  Returns the same list of dicts (tensors on the CPU)."""
    pred = results["pred"]
    num_atoms, _, edge_index = results["lig_info"][:3]
    n, node_off, edge_off = _offsets(num_atoms)
    if n.numel() != n_graphs:
        raise ValueError(f"lig_info holds {n.numel()} molecules, n_graphs = {n_graphs}")
    atom_type = pred[0].argmax(dim=-1)
    keep = atom_type < len(ATOM_TYPES)
    # index of every atom among the kept atoms of its own molecule (-1 for mask atoms)
    start = torch.as_tensor(node_off[:-1], device=keep.device)
    mol = torch.repeat_interleave(torch.arange(n_graphs, device=keep.device), n.to(keep.device))
    csum = torch.cumsum(keep.long(), 0)
    before = torch.cat([csum.new_zeros(1), csum])[start]              # kept atoms in earlier molecules
    renum = torch.where(keep, csum - 1 - before[mol], csum.new_full((), -1))
    atom_type_c, keep_c, pos_c = atom_type.to("cpu", torch.uint8), keep.cpu(), pred[1].detach().cpu()
    if include_bond:
        edge_type = pred[2].argmax(dim=-1)
        is_bond = (edge_type > 0) & (edge_type < num_bond_types)
        local = renum[edge_index]                                       # [2,E] molecule-local kept-atom indices
        ok = is_bond & (local >= 0).all(dim=0)
        edge_type_c, ok_c, local_c = edge_type.to("cpu", torch.uint8), ok.cpu(), local.cpu()
    out = []
    for i in range(n_graphs):
        a0, a1, e0, e1 = node_off[i], node_off[i + 1], edge_off[i], edge_off[i + 1]
        k = keep_c[a0:a1]
        d = {"element": [ATOM_TYPES[t] for t in atom_type_c[a0:a1][k].tolist()], "atom_pos": pos_c[a0:a1][k],
             "bond_type": None, "bond_index": None}
        if include_bond:
            m = ok_c[e0:e1]
            d["bond_type"] = edge_type_c[e0:e1][m].long()
            d["bond_index"] = local_c[:, e0:e1][:, m]
        out.append(d)
    return out
