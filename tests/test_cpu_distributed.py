"""world_size-2 gloo test of the molecule-sharded sampling plumbing: shard bounds, size-balanced
assignment and the final ragged gather of fixed-stride result records."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phoregen_b200.distributed import balanced_assignment, gather_results, pack_results, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1024, 25600):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_assignment_balances_cubic_cost():
    rng = np.random.default_rng(0)
    na = rng.integers(20, 41, size=512)
    buckets = balanced_assignment(na, 8)
    assert sorted(np.concatenate(buckets).tolist()) == list(range(512))
    load = np.array([(na[b].astype(float) ** 3).sum() for b in buckets])
    assert load.max() / load.mean() < 1.02


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    na_all = torch.tensor([5, 9, 4, 7, 6])                     # 5 molecules over 2 ranks: ragged shards
    lo, hi = shard_bounds(len(na_all), rank, world)
    na = na_all[lo:hi]
    nl, ne = int(na.sum()), int((na * (na - 1)).sum())
    g = torch.Generator().manual_seed(100 + rank)
    local = pack_results(torch.randn(nl, 3, generator=g), torch.randint(0, 12, (nl,), generator=g),
                         torch.randint(0, 6, (ne,), generator=g), na)
    got = gather_results(local, dst=0)
    torch.save(dict(local=local, got=got), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_results_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "r1.pt", weights_only=False)
    assert r1["got"] is None
    got = r0["got"]
    for key in ("num_atoms", "pos", "node_cls", "edge_cls"):
        want = torch.cat([r0["local"][key], r1["local"][key]])
        assert torch.equal(got[key], want), key
    assert got["num_atoms"].tolist() == [5, 9, 4, 7, 6]
