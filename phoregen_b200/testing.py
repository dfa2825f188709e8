"""Deterministic, torch-RNG-independent weight initialisation shared by the tests, the golden-fixture generator
and bench.py (no pretrained checkpoint ships with the reference: ckpt/README.md:3)."""
import hashlib

import numpy as np
import torch


def random_state_dict(model, seed=0):
    """Fill every trainable tensor of `model` from numpy's PCG64 stream (keys visited in sorted order) and keep
    frozen tables / buffers.  LayerNorm affines are perturbed away from (1, 0) so that they are exercised."""
    rng = np.random.default_rng(seed)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    for k in sorted(sd):
        if k not in trainable:
            continue
        shape = tuple(sd[k].shape)
        if k.endswith(".net.1.weight"):
            v = 1.0 + 0.2 * rng.standard_normal(shape)
        elif k.endswith(".net.1.bias"):
            v = 0.1 * rng.standard_normal(shape)
        elif len(shape) == 2:
            bound = 1.0 / np.sqrt(shape[1])
            v = rng.uniform(-bound, bound, shape) * 1.7
        else:
            v = rng.uniform(-0.1, 0.1, shape)
        sd[k] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return sd


def state_dict_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def grad_digest(named_grads, seed=7):
    """Per-tensor (L2 norm, projection on a seeded direction): a compact fingerprint of the 5.2 M gradient values of one
    backward pass, small enough to commit as a fixture (tests/golden/train_grads.pt)."""
    out = {}
    for k in sorted(named_grads):
        g = named_grads[k].detach().double().cpu().reshape(-1)
        r = torch.from_numpy(np.random.default_rng(seed + g.numel()).standard_normal(g.numel()))
        out[k] = (float(g.norm()), float((g * r).sum()))
    return out


MODEL_CONFIG = {   # `model:` section of reference configs/train_lig-phore.yml with phore_feat_dim += 2 (sample_all.py:41-43)
    "name": "diffusion", "num_atom_classes": 12, "num_bond_classes": 6, "lig_feat_dim": 12, "phore_feat_dim": 18,
    "hidden_dim": 128, "bond_diffusion": True, "bond_net_type": "lin", "bond_len_loss": False,
    "count_pred_type": "boundary", "loss_weight": [1, 100, 100], "count_factor": 1, "hp_emb_with_pos": True,
    "diff": {
        "num_timesteps": 1000, "time_dim": 10, "categorical_space": "discrete",
        "diff_pos": {"beta_schedule": "advance", "scale_start": 0.9999, "scale_end": 0.0001, "width": 3},
        "diff_atom": {"init_prob": "tomask", "beta_schedule": "advance", "scale_start": 0.9999, "scale_end": 0.0001, "width": 3},
        "diff_bond": {"init_prob": "absorb", "beta_schedule": "segment", "time_segment": [600, 400],
                      "segment_diff": [{"scale_start": 0.9999, "scale_end": 0.001, "width": 3},
                                       {"scale_start": 0.001, "scale_end": 0.0001, "width": 2}]},
    },
    "denoiser": {"name": "uni_node_edge", "num_blocks": 1, "num_layers": 6, "hidden_dim": 128, "n_heads": 16, "knn": 32,
                 "edge_feat_dim": 4, "num_r_gaussian": 20, "act_fn": "relu", "norm": True, "cutoff_mode": "knn",
                 "r_max": 10.0, "x2h_out_fc": False, "h_node_in_bond_net": True, "direction_match": True},
}


class PhoreStore(dict):
    """Duck-typed stand-in for a PyG HeteroData node store."""
    __getattr__ = dict.__getitem__


class PhoreData:
    """What PhoreDiff.sample needs from the reference's HeteroData: data['phore'].{x,pos,norm} and data.center."""

    def __init__(self, x, pos, norm, center=None, name="synthetic"):
        self._phore = PhoreStore(x=x, pos=pos, norm=norm)
        self.center = torch.zeros(3) if center is None else center
        self.name = name

    def __getitem__(self, key):
        if key != "phore":
            raise KeyError(key)
        return self._phore


class _NS(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class TrainBatch:
    """PyG-free stand-in for the collated training batch `compute_loss` reads (diffusion.py:249-352):
    data['ligand'].{x,pos,batch,ptr}, data['ligand','ligand'].{f_edge_attr,f_edge_index,f_edge_attr_batch,edge_index},
    data['phore'].{x,pos,norm,batch}, data.num_graphs."""

    def __init__(self, ligand, bonds, phore, num_graphs):
        self._s = {"ligand": _NS(ligand), ("ligand", "ligand"): _NS(bonds), "phore": _NS(phore)}
        self.num_graphs = num_graphs

    def __getitem__(self, key):
        return self._s[key]

    def to(self, device):
        for st in self._s.values():
            for k, v in list(st.items()):
                st[k] = v.to(device)
        return self


def training_batch_from_synthetic(b):
    """`synthetic.synthetic_batch(..., edge_order="training")` -> TrainBatch with class-index targets (the one-hot rows of the
    synthetic batch play the role of the clean molecule)."""
    n = torch.bincount(b["batch_node"])
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(n, 0)])
    ph = b["phore"]
    return TrainBatch(
        ligand=dict(x=b["h_node"].argmax(-1), pos=b["pos"].clone(), batch=b["batch_node"], ptr=ptr),
        bonds=dict(f_edge_attr=b["h_edge"].argmax(-1), f_edge_index=b["edge_index"], f_edge_attr_batch=b["batch_edge"],
                   edge_index=b["edge_index"]),
        phore=dict(x=ph["x"], pos=ph["pos"], norm=ph["norm"], batch=ph["batch"]),
        num_graphs=int(n.numel()))
