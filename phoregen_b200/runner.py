"""Many pharmacophores x many samples as ONE molecule-sharded job (BASELINE.json configs[2]).

The reference samples one pharmacophore at a time: `for pharmacophore: while finished < num_samples: model.sample(batch)`
(sample_all.py:69-94), 30 molecules per call on one GPU.  Every molecule is independent for the whole trajectory
(SURVEY.md §8(e)), so the job is a flat list of work items

    item = pharmacophore_index * num_samples + sample_index          (also the molecule's random-stream id)

dealt to the ranks by `distributed.balanced_assignment` (equal sum n^3), run in batches that mix pharmacophores
(`TrajectorySampler(phore_batch=...)`), and gathered once at the end.  Because the kernel a molecule runs on and every
random draw it consumes depend on the molecule alone (per-molecule Philox streams, engine.BatchPlan.molecule_streams),
the gathered job is bit-identical for every world size and batch size.
"""
import numpy as np
import torch

from .distributed import balanced_assignment, gather_results, pack_results


class PhoreSet:
    """P pharmacophores collated once on the device: x [Ptot,18], pos, norm, per-pharmacophore offsets and centres."""

    def __init__(self, items, device):
        self.device = torch.device(device)
        ph = [it["phore"] if not isinstance(it, dict) or "phore" in it else it for it in items]
        get = lambda p, k: p[k] if isinstance(p, dict) else getattr(p, k)
        self.count = np.array([int(get(p, "x").shape[0]) for p in ph], dtype=np.int64)
        if (self.count < 1).any():
            raise ValueError("every pharmacophore needs at least one feature")
        self.offset = np.concatenate([[0], np.cumsum(self.count)])
        cat = lambda k: torch.cat([get(p, k).float() for p in ph]).to(self.device).contiguous()
        self.x, self.pos, self.norm = cat("x"), cat("pos"), cat("norm")
        cen = [getattr(it, "center", None) for it in items]
        self.center = None if any(c is None for c in cen) else torch.stack([torch.as_tensor(c).float() for c in cen]).to(self.device)
        self.names = [getattr(it, "name", str(i)) for i, it in enumerate(items)]
        self._guidance_center = {}
        self.batch = torch.repeat_interleave(torch.arange(len(ph), device=self.device), torch.from_numpy(self.count).to(self.device))
        self._count_d = torch.from_numpy(self.count).to(self.device)
        self._offset_d = torch.from_numpy(self.offset[:-1]).to(self.device)

    def __len__(self):
        return len(self.count)

    def batch_for(self, phore_ids):
        """Pharmacophore batch (one graph per entry of `phore_ids`, repeats allowed) gathered on the device."""
        pid = torch.as_tensor(np.asarray(phore_ids, dtype=np.int64), device=self.device)
        cnt = self._count_d[pid]
        graph = torch.repeat_interleave(torch.arange(pid.numel(), device=self.device), cnt)
        first = torch.cumsum(cnt, 0) - cnt                                   # first row of each graph in the batch
        rows = self._offset_d[pid][graph] + (torch.arange(int(cnt.sum()), device=self.device) - first[graph])
        out = {"x": self.x[rows], "pos": self.pos[rows], "norm": self.norm[rows], "batch": graph}
        if self.center is not None:
            out["center"] = self.center[pid]
        if self._guidance_center:
            out["guidance_center"] = next(iter(self._guidance_center.values()))[pid]
        return out

    def with_guidance_centers(self, ex_col):
        """Per-pharmacophore non-EX centres for the guidance energy (diffusion.py:493-497), computed once per pharmacophore."""
        if ex_col not in self._guidance_center:
            from .diffusion import non_ex_centers
            self._guidance_center = {ex_col: non_ex_centers(self.x.cpu(), self.pos.cpu(), self.count, ex_col).to(self.device)}
        return self


def split_batches(num_atoms, batch_size, max_edges=None):
    """Consecutive batches of at most `batch_size` molecules and (optionally) `max_edges` directed bond edges
    (sum n(n-1): what the work space scales with)."""
    out, cur, edges = [], [], 0
    for i, n in enumerate(np.asarray(num_atoms).tolist()):
        e = n * (n - 1)
        if cur and (len(cur) >= batch_size or (max_edges is not None and edges + e > max_edges)):
            out.append(cur)
            cur, edges = [], 0
        cur.append(i)
        edges += e
    if cur:
        out.append(cur)
    return out


class SamplingJob:
    """`num_samples` molecules for each pharmacophore of `phores` (a list of PhoreData-like items, phore_io.parse_phore_file
    or testing.PhoreData).  `ligand_num_atoms`: optional [P * num_samples] atom counts (item order); by default the
    atom-count heads give every pharmacophore its interval on the device (PhoreDiff.atom_count_intervals) and one count is
    drawn per item."""

    def __init__(self, model, phores, num_samples, device, seed=2032, batch_size=1024, max_edges=None, sample_mode="uniform",
                 normal_scale=4.0, guidance=None, guidance_n_graphs=30, ligand_num_atoms=None, rank=0, world_size=1):
        self.model, self.device = model, torch.device(device)
        self.set = phores if isinstance(phores, PhoreSet) else PhoreSet(phores, device)
        self.P, self.S, self.seed = len(self.set), int(num_samples), int(seed)
        self.batch_size, self.max_edges, self.guidance = int(batch_size), max_edges, guidance
        # the reference's guidance drift scales with 1 / (graphs per sample() call) (sample_utils.py:155,165); a job fixes that
        # number (default 30: sample_all.py:24 / sample.sh:27) so the result does not depend on how the job is batched
        self.guidance_n_graphs = int(guidance_n_graphs)
        if guidance:
            self.set.with_guidance_centers(model._ex_col)
        self.rank, self.world = int(rank), int(world_size)
        n_items = self.P * self.S
        if ligand_num_atoms is None:
            # identical on every rank: same kernels, same seeded device generator (one D2H for the whole job)
            lo, hi = model.atom_count_intervals(self.set.x, self.set.pos, self.set.batch, self.P, self.device)
            gen = torch.Generator(device=self.device)
            gen.manual_seed(self.seed)
            n = model.sample_from_intervals(lo.repeat_interleave(self.S), hi.repeat_interleave(self.S), sample_mode, normal_scale, generator=gen)
            self.intervals = (lo.cpu().numpy(), hi.cpu().numpy())
            self.num_atoms = n.cpu().numpy().astype(np.int64)
        else:
            self.intervals = None
            self.num_atoms = np.asarray(torch.as_tensor(ligand_num_atoms).cpu().numpy(), dtype=np.int64)
            if self.num_atoms.shape != (n_items,):
                raise ValueError(f"ligand_num_atoms must hold {n_items} counts (pharmacophore-major)")
        self.items = balanced_assignment(self.num_atoms, self.world)[self.rank]          # ascending item ids of this rank
        self.batches = [self.items[b] for b in split_batches(self.num_atoms[self.items], self.batch_size, self.max_edges)]

    def sampler(self, items, **kw):
        from .diffusion import TrajectorySampler
        pid = items // self.S
        return TrajectorySampler(self.model, None, len(items), self.device, ligand_num_atoms=torch.from_numpy(self.num_atoms[items]),
                                 guidance=self.guidance, guidance_n_graphs=self.guidance_n_graphs, seed=self.seed,
                                 phore_batch=self.set.batch_for(pid), graph_uid=items,
                                 **{"save_traj": False, **kw})

    def run(self, num_steps=None, use_cuda_graph=True):
        """Runs this rank's batches; -> fixed-stride records of its molecules in ascending item order."""
        parts = []
        for items in self.batches:
            s = self.sampler(items, use_cuda_graph=use_cuda_graph)
            s.run(num_steps)
            rec = pack_results(s.pos + s.center_rows, s.node_cls, s.edge_cls, s.num_atoms)
            rec["item"] = torch.from_numpy(items).to(self.device)
            parts.append(rec)
            del s
        if not parts:
            z = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt, device=self.device)
            return dict(pos=z(0, 3), node_cls=z(0, dt=torch.uint8), edge_cls=z(0, dt=torch.uint8), num_atoms=z(0, dt=torch.int32), item=z(0, dt=torch.int64))
        return {k: torch.cat([p[k] for p in parts]) for k in parts[0]}


def order_by_item(rec):
    """Records of several ranks concatenated in rank order -> ascending item order (ragged per-molecule blocks)."""
    item = rec["item"].cpu()
    n = rec["num_atoms"].cpu().long()
    order = torch.argsort(item)
    a_off = torch.cumsum(n, 0) - n
    e_cnt = n * (n - 1)
    e_off = torch.cumsum(e_cnt, 0) - e_cnt

    def rows(off, cnt):
        c = cnt[order]
        first = torch.cumsum(c, 0) - c
        g = torch.repeat_interleave(torch.arange(order.numel()), c)
        return off[order][g] + (torch.arange(int(c.sum())) - first[g])
    ar, er = rows(a_off, n), rows(e_off, e_cnt)
    dev = rec["pos"].device
    return dict(pos=rec["pos"][ar.to(dev)], node_cls=rec["node_cls"][ar.to(dev)], edge_cls=rec["edge_cls"][er.to(dev)],
                num_atoms=rec["num_atoms"][order.to(dev)], item=rec["item"][order.to(dev)])


def gather_job(local, dst=0, group=None):
    """Final gather of a sharded job (the only collective): every rank's records to `dst`, put back in item order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return order_by_item(local)
    got = gather_results(local, dst=dst, group=group, extra=("item",))
    return order_by_item(got) if got is not None else None
