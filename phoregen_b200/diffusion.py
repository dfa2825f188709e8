"""Drop-in for reference models/diffusion.py `PhoreDiff` (inference tier): same constructor, attribute names,
641-key state_dict, `forward` and `sample` signatures and return layouts; the arithmetic runs in the CUDA library.

    from phoregen_b200.diffusion import PhoreDiff        # instead of `from models.diffusion import PhoreDiff`

Reference call sites served: sample_all.py:57-60,91-94 (construct, load_state_dict, sample) and
models/diffusion.py:175-246 (forward).  `compute_loss` (training tier, diffusion.py:249-352) is not part of round 1.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import _lib, schedules
from .engine import BatchPlan, PackedModel
from .modules import get_denoiser_net, get_phore_encoder


class _Cfg(dict):
    """Minimal attribute-dict so plain YAML dicts work where the reference expects an EasyDict."""

    def __init__(self, d):
        super().__init__()
        for k, v in dict(d).items():
            self[k] = _Cfg(v) if isinstance(v, dict) else ([_Cfg(i) if isinstance(i, dict) else i for i in v]
                                                         if isinstance(v, list) else v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _frozen(arr):
    return nn.Parameter(torch.from_numpy(np.asarray(arr)).float(), requires_grad=False)


class ContigousTransition(nn.Module):
    """Frozen tables of transition.py:9-26 (state_dict keys pos_transition.*)."""

    def __init__(self, betas):
        super().__init__()
        for k, v in schedules.gaussian_tables(betas).items():
            setattr(self, k, _frozen(v))


class GeneralCategoricalTransition(nn.Module):
    """Frozen tables of transition.py:178-215 (state_dict keys {node,edge}_transition.*)."""

    def __init__(self, betas, num_classes, init_prob=None):
        super().__init__()
        self.num_classes = num_classes
        tabs, prior = schedules.categorical_tables(betas, num_classes, init_prob)
        self.init_prob = prior
        self.q_mats = _frozen(tabs["q_mats"])
        self.transpopse_q_onestep_mats = _frozen(tabs["transpopse_q_onestep_mats"])

    def init_log_prob(self):
        return torch.log(torch.from_numpy(self.init_prob) + 1e-30).clamp_min(-32.0).float()


class TimeGaussianSmearing(nn.Module):
    """Buffers of common.py:34-49 (type_='linear')."""

    def __init__(self, stop, num_gaussians):
        super().__init__()
        offset = torch.linspace(0.0, float(stop), num_gaussians)
        diff = torch.diff(offset)
        diff = torch.cat([diff[:1], diff])
        self.register_buffer("coeff", -0.5 / (diff ** 2))
        self.register_buffer("offset", offset)


class _Head(nn.Sequential):
    pass


class PhoreDiff(nn.Module):
    def __init__(self, config, data_name="zinc_300", **kwargs):
        super().__init__()
        config = config if hasattr(config, "denoiser") and not isinstance(config, dict) else _Cfg(config)
        self.config, self.data_name = config, data_name
        self.num_node_types, self.num_edge_types = config.num_atom_classes, config.num_bond_classes
        self.bond_len_loss, self.bond_diffusion = config.bond_len_loss, config.bond_diffusion
        self.bond_net_type, self.count_pred_type = config.bond_net_type, config.count_pred_type
        self.max_atom, self.min_atom = 78, 4
        self.loss_weight = getattr(config, "loss_weight", [1, 100, 100])
        self.count_factor = getattr(config, "count_factor", 1)
        self.hp_emb_with_pos = getattr(config, "hp_emb_with_pos", False)
        d = config.diff
        if (self.num_node_types, self.num_edge_types, config.hidden_dim, d.time_dim) != (12, 6, 128, 10) or \
                not self.bond_diffusion or self.bond_net_type != "lin" or not self.hp_emb_with_pos or \
                d.categorical_space != "discrete" or self.count_pred_type != "boundary":
            raise NotImplementedError("phoregen_b200 kernels are compiled for the train_lig-phore.yml model section")
        self.num_timesteps, self.categorical_space = d.num_timesteps, d.categorical_space
        sched = lambda c: schedules.beta_schedule(c.beta_schedule, self.num_timesteps,
                                                  **{k: v for k, v in dict(c).items() if k not in ("beta_schedule", "init_prob")})
        self.pos_transition = ContigousTransition(sched(d.diff_pos))
        self.node_transition = GeneralCategoricalTransition(sched(d.diff_atom), 12, d.diff_atom.init_prob)
        self.edge_transition = GeneralCategoricalTransition(sched(d.diff_bond), 6, d.diff_bond.init_prob)
        self.node_embedder = nn.Linear(12, config.hidden_dim - d.time_dim, bias=False)
        self.edge_embedder = nn.Linear(6, config.hidden_dim - d.time_dim, bias=False)
        self.time_emb = nn.Sequential(TimeGaussianSmearing(self.num_timesteps, d.time_dim))
        self.phore_embedding = nn.Linear(config.phore_feat_dim, config.hidden_dim)
        if config.phore_feat_dim != 18:
            raise NotImplementedError("phore_feat_dim must be 18 at run time (16 in the YAML + 2: sample_all.py:41-43)")
        self.phore_encoder = get_phore_encoder(config.denoiser)
        self.denoiser = get_denoiser_net(config.denoiser)
        H = config.hidden_dim
        self.v_inference = _Head(nn.Linear(H, H), nn.Identity(), nn.Linear(H, 12))          # index 1 = ShiftedSoftplus (no params)
        from .modules import GaussianSmearing
        self.distance_expansion = GaussianSmearing(0.0, 5.0, num_gaussians=config.denoiser.num_r_gaussian, fix_offset=False)
        self.bond_inference = _Head(nn.Linear(H, H), nn.Identity(), nn.Linear(H, 6))
        self.atom_mlp = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 1), nn.Sigmoid())
        self.atom_mlp_1 = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 1), nn.Sigmoid())
        self._packed, self._packed_key = None, None
        self._plan, self._plan_key = None, None

    # ------------------------------------------------------------------ packing / plans
    def packed(self, device=None):
        device = torch.device(device) if device is not None else self.node_embedder.weight.device
        params = list(self.parameters())
        key = (str(device), tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if self._packed is None or self._packed_key != key:
            self._packed = PackedModel(self.state_dict(), device)
            self._packed_key = key
        return self._packed

    def _plan_for(self, batch_node, batch_phore, edge_index, n_graphs, device):
        """Plan of this batch's topology.  A cached plan is reused only when the CONTENTS agree: the per-graph atom and
        pharmacophore counts and the sortedness are recomputed on every call (one small D2H), the edge list is compared
        element-wise.  Tensor addresses are not part of the key (the caching allocator hands the same address to
        different batches)."""
        chk = torch.stack([(batch_node[1:] < batch_node[:-1]).any(), (batch_phore[1:] < batch_phore[:-1]).any()])
        counts = torch.cat([torch.bincount(batch_node, minlength=n_graphs), torch.bincount(batch_phore, minlength=n_graphs),
                            chk.long()]).cpu().numpy()
        if counts[-2] or counts[-1]:
            raise ValueError("batch vectors must be sorted by graph")
        if counts.size != 2 * n_graphs + 2:
            raise ValueError("graph ids in the batch vectors exceed the number of graphs")
        na, npn = counts[:n_graphs].astype(np.int32), counts[n_graphs:2 * n_graphs].astype(np.int32)
        pl = self._plan
        if pl is None or pl.device != torch.device(device) or not np.array_equal(pl.num_atoms, na) or not np.array_equal(pl.num_phore, npn) \
                or self._plan_edges.shape != edge_index.shape or not torch.equal(self._plan_edges, edge_index):
            self._plan = BatchPlan(na, npn, device, ref_edge_index=edge_index)
            self._plan_edges = edge_index.clone()
        return self._plan

    # ------------------------------------------------------------------ O2 (diffusion.py:148-163)
    @property
    def _ex_col(self):
        return 12 if self.data_name in ("zinc_300", "pdbbind") else 10

    def predict_atom_count(self, h_p, batch_p, _h_p, n_graphs=None, plan=None):
        """(count_l [G,1], count_u [G,1]) from the encoded pharmacophore nodes: one launch of atom_count_kernel
        (pg_atom_count) over the plan's pharmacophore segments."""
        if plan is None:
            n_graphs = int(batch_p.max().item()) + 1 if n_graphs is None else n_graphs
            plan = BatchPlan(np.full(n_graphs, 2, np.int32), torch.bincount(batch_p, minlength=n_graphs).cpu().numpy(), h_p.device, edge_order=1)
        return plan.atom_count(self.packed(h_p.device), h_p, _h_p, self._ex_col, self.min_atom, self.max_atom)

    # ------------------------------------------------------------------ forward (diffusion.py:175-246)
    @torch.no_grad()
    def forward(self, h_node_pert, pos_pert, batch_node, h_edge_pert, edge_index, batch_edge, time_step,
                h_phore, pos_phore, phore_norm, batch_phore, plan=None, h_phore_emb=None):
        dev = pos_pert.device
        n_graphs = int(time_step.numel())
        plan = plan or self._plan_for(batch_node, batch_phore, edge_index, n_graphs, dev)
        pm = self.packed(dev)
        if h_phore_emb is None:
            h_phore_emb = plan.phore_encode(pm, h_phore, pos_phore)
        v, pos, b = plan.phorediff_forward(pm, h_node_pert, pos_pert, h_edge_pert, time_step.to(torch.int64).contiguous(),
                                           h_phore_emb, pos_phore, phore_norm)
        cnt = self.predict_atom_count(h_phore_emb, batch_phore, h_phore, n_graphs, plan=plan)
        return v, pos, b, cnt

    def compute_loss(self, data, rng_device=None, provider=None):
        """The training objective (diffusion.py:249-352) -> (loss_total, loss_dict).
        With gradients enabled the loss carries an autograd graph over every trainable parameter (training.py: the
        reference formulation on torch operators, graph artefacts from the CUDA graph kernels).  Under `torch.no_grad()`
        - the reference's validation loop - the forward runs on the CUDA kernels (losses.py)."""
        if torch.is_grad_enabled():
            from . import training
            return training.compute_loss_with_grad(self, data, provider=provider, rng_device=rng_device)
        from . import losses
        return losses.compute_loss(self, data, rng_device=rng_device)

    # ------------------------------------------------------------------ D2 (diffusion.py:356-387)
    @torch.no_grad()
    def atom_count_intervals(self, phore_x, phore_pos, phore_batch, n_graphs, device):
        """Integer atom-count interval [lo, hi] of every graph of a pharmacophore batch (diffusion.py:356-380), on the
        device: phore encoder + atom_count_kernel, no host synchronisation.  -> (lo [G] int32, hi [G] int32)."""
        device = torch.device(device)
        num_phore = torch.bincount(phore_batch, minlength=n_graphs).cpu().numpy().astype(np.int32)
        plan = BatchPlan(np.full(n_graphs, 2, np.int32), num_phore, device, edge_order=1)   # ligand side unused by the encoder
        pm = self.packed(device)
        x, pos = phore_x.to(device).float().contiguous(), phore_pos.to(device).float().contiguous()
        h_p = plan.phore_encode(pm, x, pos)
        _, _, lo, hi = plan.atom_count(pm, h_p, x, self._ex_col, self.min_atom, self.max_atom, intervals=True)
        return lo, hi

    @staticmethod
    def sample_from_intervals(lo, hi, mode="uniform", scale=4.0, generator=None):
        """utils/sample_utils.py:28-37 for per-graph intervals, drawn on the device of `lo` (one draw per graph)."""
        lo_f, hi_f = lo.float(), hi.float()
        if mode == "uniform":                                   # randint(lo, hi + 1)
            u = torch.rand(lo.shape, device=lo.device, generator=generator)
            return (lo + torch.floor(u * (hi_f - lo_f + 1.0)).to(lo.dtype)).clamp(max=hi)
        if mode == "normal":
            z = torch.randn(lo.shape, device=lo.device, generator=generator)
            n = (lo_f + hi_f) / 2 + z * (hi_f - lo_f) / scale
            return torch.minimum(torch.maximum(n, lo_f), hi_f).round().to(lo.dtype)
        raise NotImplementedError(f"The sample nodes mode {mode} is not implemented.")

    @torch.no_grad()
    def sample_nodes(self, data, batch_size, device, sample_mode="uniform", normal_scale=4.0, return_interval=False):
        """diffusion.py:356-387: `batch_size` atom counts for ONE pharmacophore.  The interval comes from the device
        kernels; the draws use the CPU generator like the reference's sample_from_interval (sample_utils.py:28-37), so
        a seeded run draws the same counts as the reference."""
        ph = data["phore"]
        x = ph.x.to(device).float()
        lo, hi = self.atom_count_intervals(x, ph.pos, torch.zeros(x.shape[0], dtype=torch.long, device=device), 1, device)
        lo, hi = int(lo.item()), int(hi.item())
        if return_interval:
            return lo, hi
        if sample_mode == "uniform":
            n = torch.randint(lo, hi + 1, (batch_size,))
        elif sample_mode == "normal":
            n = torch.normal((lo + hi) / 2, (hi - lo) / normal_scale, (batch_size,)).clamp(lo, hi).round().int()
        else:
            raise NotImplementedError(f"The sample nodes mode {sample_mode} is not implemented.")
        return n.to(device)

    # ------------------------------------------------------------------ D1 (diffusion.py:390-525)
    @torch.no_grad()
    def sample(self, data, n_graphs, device, pos_guidance_opt=None, sample_mode="uniform", normal_scale=4.0,
               ligand_num_atoms=None, save_traj=True, seed=None, use_cuda_graph=True, num_steps=None, traj_layout="compact",
               phore_batch=None, **kwargs):
        """Reverse diffusion for `n_graphs` copies of one pharmacophore.  Returns the reference's dict
        {'pred': [logits_node, pos+center, logits_edge], 'traj': [node, pos, edge], 'lig_info': [...]}.
        Extensions (keyword-only in spirit): `ligand_num_atoms` bypasses the atom-count head, `save_traj=False`
        keeps only the final state (time dim 1), `seed` fixes the Philox stream, `num_steps` truncates the loop
        (benchmarks / tests)."""
        device = torch.device(device)
        sampler = TrajectorySampler(self, data, n_graphs, device, ligand_num_atoms=ligand_num_atoms,
                                    sample_mode=sample_mode, normal_scale=normal_scale, guidance=pos_guidance_opt,
                                    save_traj=save_traj, seed=seed, use_cuda_graph=use_cuda_graph, phore_batch=phore_batch)
        sampler.run(num_steps)
        return sampler.results(traj_layout)


def non_ex_centers(x, pos, num_phore, ex_col):
    """diffusion.py:493-497 per graph: mean position of the features that are not exclusion spheres, [G,3].  Computed graph by
    graph on the host (a device scatter-add would make the centre depend on atomics' order and on the batch composition;
    the sharded job needs it to depend on the pharmacophore alone).  nan for a graph without such features, like the
    reference's mean over an empty selection."""
    out, o = [], 0
    for p in np.asarray(num_phore).tolist():
        xs, ps = x[o:o + p], pos[o:o + p]
        out.append(ps[xs[:, ex_col] != 1].mean(0))
        o += p
    return torch.stack(out).float().contiguous()


class TrajectorySampler:
    """State + one captured CUDA graph of a reverse step (forward + categorical/Gaussian posterior update).

    All shapes are static across the trajectory (bond/triplet topology is fixed and the kNN edge count is
    N*min(32,N-1) per graph whatever the coordinates), the time step and the RNG counter live in device scalars,
    so the whole step is captured once and replayed `num_timesteps` times."""

    def __init__(self, model, data, n_graphs, device, ligand_num_atoms=None, sample_mode="uniform", normal_scale=4.0,
                 guidance=None, save_traj=True, seed=None, use_cuda_graph=True, phore_batch=None, graph_uid=None,
                 guidance_n_graphs=0):
        """`data` is one pharmacophore replicated `n_graphs` times (the reference's sample()); alternatively
        `phore_batch` = dict(x [P,18], pos [P,3], norm [P,3], batch [P]) supplies a different pharmacophore per
        molecule (something the reference's sample() cannot do: sample_all.py:69-94 handles one at a time).
        Every random draw of molecule g (initial state, Gumbel noise, position noise) comes from its own Philox stream
        keyed by (seed, graph_uid[g]) - default uid = g - so a molecule's trajectory does not depend on the batch it is in
        (engine.BatchPlan.molecule_streams); a sharded job passes job-wide ids."""
        self.m, self.device, self.G = model, device, n_graphs
        self.T = model.num_timesteps
        pm = self.pm = model.packed(device)
        if phore_batch is None:
            ph = data["phore"]
            px, ppos, pnorm = ph.x.to(device).float(), ph.pos.to(device).float(), ph.norm.to(device).float()
            if ligand_num_atoms is None:
                ligand_num_atoms = model.sample_nodes(data, n_graphs, device, sample_mode, normal_scale)
            num_phore = np.full(n_graphs, px.shape[0], dtype=np.int32)
            center = getattr(data, "center", None)
            # Batch.from_data_list([data.clone()] * n_graphs)  (diffusion.py:399)
            self.px, self.ppos, self.pnorm = px.repeat(n_graphs, 1), ppos.repeat(n_graphs, 1).contiguous(), pnorm.repeat(n_graphs, 1).contiguous()
            single_x, single_pos = px, ppos
        else:
            self.px = phore_batch["x"].to(device).float().contiguous()
            self.ppos = phore_batch["pos"].to(device).float().contiguous()
            self.pnorm = phore_batch["norm"].to(device).float().contiguous()
            pb = phore_batch["batch"].to(device)
            cnt = torch.bincount(pb, minlength=n_graphs)
            bad = torch.stack([(pb[1:] < pb[:-1]).any(), (cnt == 0).any(), torch.as_tensor(cnt.numel() != n_graphs, device=device)])
            num_phore = torch.cat([cnt, bad.long()]).cpu().numpy()
            if num_phore[-3:].any():
                raise ValueError("phore_batch['batch'] must be sorted by graph, with at least one feature for each of the n_graphs graphs")
            num_phore = num_phore[:-3].astype(np.int32)
            if ligand_num_atoms is None:        # D2 for a batch of pharmacophores: intervals and draws on the device
                gen = None
                if seed is not None:
                    gen = torch.Generator(device=device)
                    gen.manual_seed(int(seed) ^ 0x5DEECE66D)
                lo, hi = model.atom_count_intervals(self.px, self.ppos, pb, n_graphs, device)
                ligand_num_atoms = model.sample_from_intervals(lo, hi, sample_mode, normal_scale, generator=gen)
            center = phore_batch.get("center")    # [G,3] per-graph centres (collate_phores) or None: positions stay centred
            single_x = single_pos = None
        self.num_atoms = torch.as_tensor(ligand_num_atoms).to(device)
        na = self.num_atoms.cpu().numpy().astype(np.int32)
        self.center = (torch.zeros(3, device=device) if center is None else torch.as_tensor(center).to(device).float()).contiguous()
        if self.center.dim() == 2 and self.center.shape != (n_graphs, 3):
            raise ValueError("per-graph centres must be [n_graphs, 3]")
        plan = self.plan = BatchPlan(na, num_phore, device, edge_order=0)
        self.h_phore_emb = plan.phore_encode(pm, self.px, self.ppos)      # step-invariant (SURVEY.md §8(d) reduction 4)
        self.batch_node = torch.repeat_interleave(torch.arange(n_graphs, device=device), self.num_atoms.long())
        self.edge_index, self.edge_batch = plan.bond_edges()                # G1 on device
        Nl, Eb = plan.Nl, plan.Eb
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        self.streams = plan.molecule_streams(self.seed, graph_uid)
        # initial state (diffusion.py:406-408; transition.py:65-69,331-339) from the init kernels (pg_position_init / pg_sample_init)
        self.center_rows = self.center[self.batch_node] if self.center.dim() == 2 else self.center     # [Nl,3] or [3]
        self.pos = plan.position_init(center=self.center, seed=self.seed, streams=self.streams)
        self.h_node, self.node_cls, self.log_node = plan.sample_init("node", model.node_transition.init_log_prob(), seed=self.seed, streams=self.streams)
        self.h_edge, self.edge_cls, self.log_edge = plan.sample_init("edge", model.edge_transition.init_log_prob(), seed=self.seed, streams=self.streams)
        self.pred = (torch.empty(Nl, 12, device=device), torch.empty(Nl, 3, device=device), torch.empty(Eb, 6, device=device))
        self.time_step = torch.full((n_graphs,), self.T - 1, dtype=torch.int64, device=device)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=device)
        self.guidance = guidance
        # the guidance energies are means over the call's n_graphs (reference quirk: the drift scales with 1 / batch size,
        # sample_utils.py:155,165); 0 keeps that, a sharded job passes one constant for all its batches
        self.guidance_n_graphs = int(guidance_n_graphs)
        self.grad = torch.zeros(Nl, 3, device=device) if guidance else None
        self.phore_center = None
        if guidance:
            col = model._ex_col
            if single_x is not None:
                self.phore_center = single_pos[single_x[:, col] != 1].mean(0).contiguous()   # diffusion.py:493-497
            elif phore_batch.get("guidance_center") is not None:     # precomputed per pharmacophore (runner.PhoreSet)
                self.phore_center = phore_batch["guidance_center"].to(device).float().contiguous()
            else:   # one pharmacophore per graph: centre of each graph's non-EX features, once per batch
                self.phore_center = non_ex_centers(self.px.cpu(), self.ppos.cpu(), num_phore, col).to(device)
        self.save_traj = save_traj
        if save_traj:
            try:
                self.traj_node = torch.zeros(self.T + 1, Nl, dtype=torch.uint8, device=device)
                self.traj_edge = torch.zeros(self.T + 1, Eb, dtype=torch.uint8, device=device)
                self.traj_pos = torch.zeros(self.T + 1, Nl, 3, device=device)
            except torch.OutOfMemoryError as e:
                raise RuntimeError(f"CUDA out of memory allocating trajectory buffers: {e}") from e
            self.traj_node[0], self.traj_edge[0], self.traj_pos[0] = self.node_cls.to(torch.uint8), self.edge_cls.to(torch.uint8), self.pos
        else:
            self.traj_node = self.traj_edge = self.traj_pos = None
        self.use_cuda_graph = use_cuda_graph
        self.graph = None
        self.steps_done = 0

    def _step(self):
        """Loop body of diffusion.py:432-517."""
        plan, pm = self.plan, self.pm
        plan.phorediff_forward(pm, self.h_node, self.pos, self.h_edge, self.time_step, self.h_phore_emb, self.ppos,
                               self.pnorm, out=self.pred)
        plan.categorical_step(pm, "node", self.pred[0], self.log_node, self.time_step, seed=self.seed, step_counter=self.step_counter,
                              onehot=self.h_node, cls=self.node_cls, traj=self.traj_node, streams=self.streams)
        plan.categorical_step(pm, "edge", self.pred[2], self.log_edge, self.time_step, seed=self.seed, step_counter=self.step_counter,
                              onehot=self.h_edge, cls=self.edge_cls, traj=self.traj_edge, streams=self.streams)
        if self.guidance:
            plan.guidance_grad(self.pos, self.edge_cls, self.guidance, self.phore_center, out=self.grad, norm_graphs=self.guidance_n_graphs)
        plan.position_step(pm, self.pos, self.pred[1], self.time_step, energy_grad=self.grad, seed=self.seed,
                           step_counter=self.step_counter, out=self.pos, traj=self.traj_pos, center=self.center, streams=self.streams)
        self.time_step.sub_(1)
        self.step_counter.add_(1)

    def run(self, num_steps=None):
        """Advance the trajectory; returns the number of steps actually executed (at most the steps that are left)."""
        n = self.T - self.steps_done if num_steps is None else min(num_steps, self.T - self.steps_done)
        if n <= 0:
            return 0
        with torch.cuda.device(self.device):      # capture and replay on the plan's GPU whatever device is current in the caller
            self._run(n)
        self.steps_done += n
        return n

    def _run(self, n):
        if not self.use_cuda_graph:
            for _ in range(n):
                self._step()
        else:
            if self.graph is None:
                # warm-up outside capture (lazy module loading, cudaFuncSetAttribute), then roll the counters back
                self._snapshot = [t.clone() for t in self._state_tensors()]
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._step()
                torch.cuda.current_stream(self.device).wait_stream(s)
                for t, c in zip(self._state_tensors(), self._snapshot):
                    t.copy_(c)
                self._snapshot = None
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step()
            for _ in range(n):
                self.graph.replay()

    def _state_tensors(self):
        ts = [self.h_node, self.pos, self.h_edge, self.log_node, self.log_edge, self.node_cls, self.edge_cls,
              self.time_step, self.step_counter]
        if self.save_traj:
            ts += [self.traj_node[1], self.traj_edge[1], self.traj_pos[1]]
        return ts

    def results(self, layout="compact"):
        """The reference's result dict (diffusion.py:519-525).  `layout`:
          "compact"   (default) the categorical trajectories stay class indices: `traj` = [ClassTrajectory(node),
                      pos [T+1,Nl,3] f32, ClassTrajectory(edge)].  A ClassTrajectory indexes like the reference's one-hot
                      tensor ([:, mask], [-1], .cpu(), .shape) and expands to one-hot f32 only for the slice that is read,
                      so unbatch_data (sample_utils.py:57-93) consumes it unchanged while the device never holds the
                      [T+1,E_b,6] f32 tensor (21 GB at configs[1]);
          "reference" the dense one-hot f32 tensors of diffusion.py:418-426 (small batches / tests)."""
        pred_pos = self.pred[1] + self.center_rows                             # reference quirk 3 (diffusion.py:519)
        if self.save_traj:
            pos_traj = self.traj_pos                                           # slot 0 is the un-centred init (diffusion.py:425)
            if layout == "reference":
                node_traj = F.one_hot(self.traj_node.long(), 12).float()
                edge_traj = F.one_hot(self.traj_edge.long(), 6).float()
            else:
                from .results import ClassTrajectory
                node_traj, edge_traj = ClassTrajectory(self.traj_node, 12), ClassTrajectory(self.traj_edge, 6)
        else:
            node_traj, pos_traj, edge_traj = self.h_node[None], (self.pos + self.center_rows)[None], self.h_edge[None]
        return {"pred": [self.pred[0], pred_pos, self.pred[2]],
                "traj": [node_traj, pos_traj, edge_traj],
                "lig_info": [self.num_atoms, self.batch_node, self.edge_index, self.edge_batch]}
