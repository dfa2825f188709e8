"""Molecule-sharded sampling across ranks: every molecule is independent for the whole
trajectory, so ranks run disjoint contiguous blocks with NO collective inside the loop and one final gather of
fixed-stride result records.  Uses torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world_size):
    """Contiguous near-equal blocks: the first `n_items % world_size` ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_assignment(num_atoms, world_size):
    """Static assignment of molecules to ranks balancing sum(n^3) (the triplet layer dominates the cost):
    sort by size, deal out greedily to the least-loaded rank.  Returns a list of index arrays (one per rank)."""
    num_atoms = np.asarray(num_atoms, dtype=np.int64)
    order = np.argsort(-num_atoms, kind="stable")
    load = np.zeros(world_size)
    buckets = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        buckets[r].append(int(i))
        load[r] += float(num_atoms[i]) ** 3
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def pack_results(pos, node_cls, edge_cls, num_atoms):
    """Fixed-stride wire records of one rank: per atom 3 x f32 + u8 class, per directed edge u8 class."""
    return dict(pos=pos.contiguous().float(), node_cls=node_cls.to(torch.uint8).contiguous(),
                edge_cls=edge_cls.to(torch.uint8).contiguous(), num_atoms=num_atoms.to(torch.int32).contiguous())


def gather_results(local, dst=0, group=None, extra=()):
    """Final gather to rank `dst`: sizes first, then one padded all_gather per field (ragged across ranks).
    `extra`: names of additional per-molecule fields of `local` (e.g. the job-wide item ids of runner.SamplingJob).
    Returns the concatenated dict on `dst` (None elsewhere)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local["pos"].device
    sizes = torch.tensor([local["num_atoms"].numel(), local["pos"].shape[0], local["edge_cls"].numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()
    out = {}
    for key, col, tail in (("num_atoms", 0, ()), ("pos", 1, (3,)), ("node_cls", 1, ()), ("edge_cls", 2, ())) + tuple((k, 0, ()) for k in extra):
        mx = int(all_sizes[:, col].max())
        buf = torch.zeros((mx,) + tail, dtype=local[key].dtype, device=dev)
        buf[: local[key].shape[0]] = local[key]
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        if rank == dst:
            out[key] = torch.cat([p[: int(all_sizes[r, col])] for r, p in enumerate(parts)])
    return out if rank == dst else None
