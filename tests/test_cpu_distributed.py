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


# ---------------------------------------------------------------- sharded job plumbing (runner.py), host side only
def test_split_batches_and_order_by_item():
    from phoregen_b200.runner import order_by_item, split_batches
    na = np.array([30, 30, 80, 20, 25, 40, 40])
    assert split_batches(na, 3) == [[0, 1, 2], [3, 4, 5], [6]]
    assert split_batches(na, 100, max_edges=2 * 870) == [[0, 1], [2], [3, 4], [5], [6]]
    # two "ranks" with interleaved item ids -> ascending item order, ragged blocks moved as a whole
    g = torch.Generator().manual_seed(0)
    n = torch.tensor([3, 5, 2, 4], dtype=torch.int32)
    item = torch.tensor([4, 9, 1, 6])
    nl, ne = int(n.sum()), int((n * (n - 1)).sum())
    rec = dict(pos=torch.randn(nl, 3, generator=g), node_cls=torch.randint(0, 12, (nl,), dtype=torch.uint8, generator=g),
               edge_cls=torch.randint(0, 6, (ne,), dtype=torch.uint8, generator=g), num_atoms=n, item=item)
    out = order_by_item(rec)
    assert out["item"].tolist() == [1, 4, 6, 9] and out["num_atoms"].tolist() == [2, 3, 4, 5]
    a_off = [0, 3, 8, 10]; e_off = [0, 6, 26, 28]
    want_pos = torch.cat([rec["pos"][a_off[k]:a_off[k] + int(n[k])] for k in (2, 0, 3, 1)])
    want_edge = torch.cat([rec["edge_cls"][e_off[k]:e_off[k] + int(n[k] * (n[k] - 1))] for k in (2, 0, 3, 1)])
    assert torch.equal(out["pos"], want_pos) and torch.equal(out["edge_cls"], want_edge)


def _job_worker(rank, world, port, out_dir):
    """Each rank fabricates the records of ITS items of a 2-rank job (what SamplingJob.run returns) and gathers them."""
    from phoregen_b200.runner import gather_job
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    na_all = np.array([5, 9, 4, 7, 6, 8, 3])
    mine = balanced_assignment(na_all, world)[rank]
    recs = []
    for it in mine:
        g = torch.Generator().manual_seed(1000 + int(it))           # a molecule's record depends on its item id only
        n = int(na_all[it])
        recs.append(dict(pos=torch.randn(n, 3, generator=g), node_cls=torch.randint(0, 12, (n,), generator=g).to(torch.uint8),
                         edge_cls=torch.randint(0, 6, (n * (n - 1),), generator=g).to(torch.uint8),
                         num_atoms=torch.tensor([n], dtype=torch.int32), item=torch.tensor([int(it)])))
    local = {k: torch.cat([r[k] for r in recs]) for k in recs[0]}
    got = gather_job(local, dst=0)
    torch.save(dict(got=got, mine=mine), os.path.join(out_dir, f"job{rank}.pt"))
    dist.destroy_process_group()


def test_gather_job_world_size_2_equals_single_rank(tmp_path):
    from phoregen_b200.runner import order_by_item
    port = _free_port()
    mp.spawn(_job_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "job0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "job1.pt", weights_only=False)
    assert r1["got"] is None and sorted(np.concatenate([r0["mine"], r1["mine"]]).tolist()) == list(range(7))
    # the same job on one rank
    na_all = np.array([5, 9, 4, 7, 6, 8, 3])
    recs = []
    for it in range(7):
        g = torch.Generator().manual_seed(1000 + it)
        n = int(na_all[it])
        recs.append(dict(pos=torch.randn(n, 3, generator=g), node_cls=torch.randint(0, 12, (n,), generator=g).to(torch.uint8),
                         edge_cls=torch.randint(0, 6, (n * (n - 1),), generator=g).to(torch.uint8),
                         num_atoms=torch.tensor([n], dtype=torch.int32), item=torch.tensor([it])))
    single = order_by_item({k: torch.cat([r[k] for r in recs]) for k in recs[0]})
    for k in single:
        assert torch.equal(r0["got"][k], single[k]), k
