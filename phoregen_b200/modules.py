"""Parameter-holding mirrors of the reference's denoiser modules.

These classes reproduce the MODULE TREE of reference models/uni_denoiser.py and models/common.py (same
attribute names, parameter shapes and buffers), so that `state_dict()` / `load_state_dict(strict=True)` and
`utils/training_utils.py:18-26` (`freeze_parameters`, which walks `denoiser.base_block[i].pos_layer_with_*`)
work unchanged.  They hold parameters only: the arithmetic runs in the CUDA library (`engine.py`), never in
PyTorch.  Factory functions mirror models/__init__.py:5-35.
"""
import numpy as np
import torch
from torch import nn

from . import _lib
from .engine import BatchPlan, PackedModel

SMEAR_OFFSETS = [0, 1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10]


class _NoTorchForward(nn.Module):
    def forward(self, *a, **k):
        raise _lib.PhoreGenLibraryError(
            f"{type(self).__name__} holds parameters only; the computation runs in libphoregen_b200 "
            "(call the owning denoiser / PhoreDiff)")


class GaussianSmearing(_NoTorchForward):
    """common.py:11-31 — only the `offset` buffer matters for the state_dict."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50, fix_offset=True):
        super().__init__()
        off = torch.tensor(SMEAR_OFFSETS, dtype=torch.float32) if fix_offset else torch.linspace(start, stop, num_gaussians)
        self.register_buffer("offset", off)


class AngularEncoding(_NoTorchForward):
    """common.py:67-87."""

    def __init__(self, num_funcs=3):
        super().__init__()
        self.register_buffer("freq_bands", torch.tensor(
            [float(i + 1) for i in range(num_funcs)] + [1.0 / (i + 1) for i in range(num_funcs)]))


class MLP(_NoTorchForward):
    """common.py:99-119 with num_layer=2, norm=True, act_fn='relu': net.0 Linear, net.1 LayerNorm, net.3 Linear."""

    def __init__(self, in_dim, out_dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                 nn.Linear(hidden_dim, out_dim))


class NodeUpdateLayer(_NoTorchForward):
    """uni_denoiser.py:13-38 (out_fc=False)."""

    def __init__(self, input_dim, hidden_dim, output_dim, n_heads, edge_feat_dim, direction_match=False, out_fc=False):
        super().__init__()
        if out_fc:
            raise NotImplementedError("x2h_out_fc=True is not used by the shipped configs and is not implemented")
        kv = input_dim * 2 + edge_feat_dim + (9 if direction_match else 0)
        self.hk_func = MLP(kv, output_dim, hidden_dim)
        self.hv_func = MLP(kv, output_dim, hidden_dim)
        self.hq_func = MLP(input_dim, output_dim, hidden_dim)


class BondUpdateLayer(_NoTorchForward):
    """uni_denoiser.py:75-99 (include_h_node=True)."""

    def __init__(self, input_dim, hidden_dim, output_dim, n_heads):
        super().__init__()
        self.distance_expansion = GaussianSmearing()
        self.angle_expansion = AngularEncoding()
        kv = input_dim + 40 + 13 + 2 * input_dim
        self.hk_func = MLP(kv, output_dim, hidden_dim)
        self.hv_func = MLP(kv, output_dim, hidden_dim)
        self.hq_func = MLP(2 * input_dim, output_dim, hidden_dim)


class PosUpdateLayer(_NoTorchForward):
    """uni_denoiser.py:168-185."""

    def __init__(self, input_dim, hidden_dim, output_dim, n_heads, edge_feat_dim, direction_match=False):
        super().__init__()
        kv = input_dim * 2 + edge_feat_dim + (9 if direction_match else 0)
        self.xk_func = MLP(kv, output_dim, hidden_dim)
        self.xv_func = MLP(kv, n_heads, hidden_dim)
        self.xq_func = MLP(input_dim, output_dim, hidden_dim)


class AttentionLayerO2TwoUpdateNodeGeneral(_NoTorchForward):
    """uni_denoiser.py:212-258."""

    def __init__(self, hidden_dim, n_heads, num_r_gaussian, edge_feat_dim):
        super().__init__()
        self.distance_expansion = GaussianSmearing()
        self.lin_node = nn.Linear(hidden_dim, hidden_dim)
        ef = num_r_gaussian * edge_feat_dim + edge_feat_dim
        self.node_layer_with_edge = NodeUpdateLayer(hidden_dim, hidden_dim, hidden_dim, n_heads, ef, direction_match=True)
        self.node_layer_with_bond = NodeUpdateLayer(hidden_dim, hidden_dim, hidden_dim, n_heads, hidden_dim)
        self.bond_layer = BondUpdateLayer(hidden_dim, hidden_dim, hidden_dim, n_heads)
        self.pos_layer_with_edge = PosUpdateLayer(hidden_dim, hidden_dim, hidden_dim, n_heads, ef, direction_match=True)
        self.pos_layer_with_bond = PosUpdateLayer(hidden_dim, hidden_dim, hidden_dim, n_heads, hidden_dim)
        self.dire_embedding = nn.Linear(3, 9)


def _check_arch(num_blocks, num_layers, hidden_dim, n_heads, k, edge_feat_dim, num_r_gaussian, act_fn, norm, cutoff_mode,
                x2h_out_fc, h_node_in_bond_net, direction_match):
    """The CUDA kernels are specialised for the train_lig-phore.yml architecture."""
    want = dict(num_blocks=1, num_layers=6, hidden_dim=128, n_heads=16, k=32, edge_feat_dim=4, num_r_gaussian=20,
                act_fn="relu", norm=True, cutoff_mode="knn", x2h_out_fc=False, h_node_in_bond_net=True, direction_match=True)
    got = dict(num_blocks=num_blocks, num_layers=num_layers, hidden_dim=hidden_dim, n_heads=n_heads, k=k,
               edge_feat_dim=edge_feat_dim, num_r_gaussian=num_r_gaussian, act_fn=act_fn, norm=norm, cutoff_mode=cutoff_mode,
               x2h_out_fc=x2h_out_fc, h_node_in_bond_net=h_node_in_bond_net, direction_match=direction_match)
    bad = {k_: (got[k_], want[k_]) for k_ in want if got[k_] != want[k_]}
    if bad:
        raise NotImplementedError(f"phoregen_b200 kernels are compiled for the train_lig-phore.yml denoiser; unsupported: {bad}")


def topology_from_context(batch, mask_ligand):
    """Per-graph (num_phore, num_atoms) from compose_context's outputs (common.py:180-208): `batch` sorted, and inside
    each graph the pharmacophore rows precede the ligand rows."""
    b = batch.detach().cpu().numpy()
    m = mask_ligand.detach().cpu().numpy().astype(bool)
    if b.size == 0 or np.any(np.diff(b) < 0):
        raise ValueError("batch must be sorted and non-empty")
    G = int(b[-1]) + 1
    n = np.bincount(b[m], minlength=G)
    p = np.bincount(b[~m], minlength=G)
    start = np.concatenate([[0], np.cumsum(n + p)])[:-1]
    pos_in_graph = np.arange(b.size) - start[b]
    if np.any(m != (pos_in_graph >= p[b])):
        raise ValueError("context rows must be ordered [pharmacophore..., ligand...] inside every graph (compose_context order)")
    return p.astype(np.int32), n.astype(np.int32)


class UniTransformerO2TwoUpdateGeneralBond(nn.Module):
    """Drop-in for reference uni_denoiser.py:301-430: same constructor keywords, parameters and forward signature;
    the forward runs pg_denoiser_forward."""

    def __init__(self, num_blocks, num_layers, hidden_dim, n_heads=1, k=32, num_bond_classes=1, num_r_gaussian=50,
                 edge_feat_dim=0, act_fn="relu", norm=True, cutoff_mode="radius", use_global_ew=True, r_max=10.0,
                 x2h_out_fc=True, h_node_in_bond_net=False, direction_match=False):
        super().__init__()
        _check_arch(num_blocks, num_layers, hidden_dim, n_heads, k, edge_feat_dim, num_r_gaussian, act_fn, norm, cutoff_mode,
                    x2h_out_fc, h_node_in_bond_net, direction_match)
        self.num_blocks, self.num_layers, self.hidden_dim, self.n_heads, self.k = num_blocks, num_layers, hidden_dim, n_heads, k
        self.num_r_gaussian, self.edge_feat_dim, self.cutoff_mode, self.r_max = num_r_gaussian, edge_feat_dim, cutoff_mode, r_max
        self.use_global_ew = use_global_ew
        self.distance_expansion = GaussianSmearing()
        self.edge_pred_layer = MLP(num_r_gaussian, 1, hidden_dim)
        self.base_block = nn.ModuleList(
            [AttentionLayerO2TwoUpdateNodeGeneral(hidden_dim, n_heads, num_r_gaussian, edge_feat_dim) for _ in range(num_layers)])
        self._packed = None
        self._packed_key = None
        self._plan = None
        self._plan_key = None

    # -- weight packing for stand-alone use (inside PhoreDiff the owner packs the whole model once)
    def _standalone_state_dict(self):
        sd = {"denoiser." + k: v for k, v in self.state_dict().items()}
        z = lambda *s: torch.zeros(*s)
        sd.update({"node_embedder.weight": z(118, 12), "edge_embedder.weight": z(118, 6), "time_emb.0.coeff": z(10),
                   "time_emb.0.offset": z(10), "phore_embedding.weight": z(128, 18), "phore_embedding.bias": z(128),
                   "v_inference.0.weight": z(128, 128), "v_inference.0.bias": z(128), "v_inference.2.weight": z(12, 128),
                   "v_inference.2.bias": z(12), "bond_inference.0.weight": z(128, 128), "bond_inference.0.bias": z(128),
                   "bond_inference.2.weight": z(6, 128), "bond_inference.2.bias": z(6)})
        for fn, din in (("hk_func", 257), ("hv_func", 257), ("hq_func", 128)):
            p = f"phore_encoder.{fn}.net."
            sd.update({p + "0.weight": z(128, din), p + "0.bias": z(128), p + "1.weight": z(128), p + "1.bias": z(128),
                       p + "3.weight": z(128, 128), p + "3.bias": z(128)})
        return sd

    def _get_packed(self, device):
        key = (str(device), tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._packed is None or self._packed_key != key:
            self._packed = PackedModel(self._standalone_state_dict(), device)
            self._packed_key = key
        return self._packed

    def forward(self, h, x, group_idx, bond_index, h_bond, mask_ligand, mask_ligand_atom, batch, phore_norm=None,
                return_all=False, packed=None, plan=None):
        if group_idx is not None:
            raise NotImplementedError("group_idx is always None in the reference (diffusion.py:213)")
        if torch.is_grad_enabled() and any(t.requires_grad for t in (h, x, h_bond)):
            raise NotImplementedError("the stand-alone denoiser module is forward only; gradients flow through PhoreDiff.compute_loss")
        dev = h.device
        if plan is None:
            # The cached plan is reused only if the CONTENTS agree: per-graph counts (recomputed on every call: two small
            # D2H copies) and the bond index.  Addresses are not a key - the caching allocator reuses them across batches.
            num_phore, num_atoms = topology_from_context(batch, mask_ligand)
            pl = self._plan
            if pl is None or pl.device != dev or not np.array_equal(pl.num_atoms, num_atoms) or not np.array_equal(pl.num_phore, num_phore) \
                    or self._plan_bond_index.shape != bond_index.shape or not torch.equal(self._plan_bond_index, bond_index):
                lig_rows = mask_ligand.nonzero()[:, 0]
                ctx_to_lig = torch.full((h.shape[0],), -1, dtype=torch.int64, device=dev)
                ctx_to_lig[lig_rows] = torch.arange(lig_rows.numel(), device=dev)
                self._plan = BatchPlan(num_atoms, num_phore, dev, ref_edge_index=ctx_to_lig[bond_index])
                self._plan_bond_index = bond_index.clone()
            plan = self._plan
        packed = packed or self._get_packed(dev)
        ho, xo, bo = plan.denoiser_forward(packed, h, x, h_bond, phore_norm)
        out = {"x": xo, "h": ho, "h_bond": bo}
        if return_all:      # uni_denoiser.py:398-430 with num_blocks = 1: the inputs, then the state after the block
            out.update({"all_x": [x, xo], "all_h": [h, ho], "all_h_bond": [h_bond, bo]})
        return out


def get_denoiser_net(config):
    """models/__init__.py:5-26."""
    if config.name != "uni_node_edge":
        raise NotImplementedError(f"Denoiser: `{config.name}` is not implemented")
    return UniTransformerO2TwoUpdateGeneralBond(
        num_blocks=config.num_blocks, num_layers=config.num_layers, hidden_dim=config.hidden_dim, n_heads=config.n_heads,
        k=config.knn, edge_feat_dim=config.edge_feat_dim, num_r_gaussian=config.num_r_gaussian, act_fn=config.act_fn,
        norm=config.norm, cutoff_mode=config.cutoff_mode, r_max=config.r_max, x2h_out_fc=config.x2h_out_fc,
        h_node_in_bond_net=config.h_node_in_bond_net, direction_match=getattr(config, "direction_match", False))


def get_phore_encoder(config):
    """models/__init__.py:29-35."""
    return NodeUpdateLayer(config.hidden_dim, config.hidden_dim, config.hidden_dim, n_heads=config.n_heads, edge_feat_dim=1,
                           out_fc=config.x2h_out_fc)
