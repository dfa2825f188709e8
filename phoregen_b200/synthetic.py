"""Seeded synthetic ligand + pharmacophore batches in the reference's tensor layout (shapes of
BASELINE.json configs[1..4]).  Pure numpy RNG (PCG64) so that the same seed gives the same batch on every machine;
used by the tests, the golden-fixture generator and bench.py.  No datasets ship with the reference's hot path and
there is no network, so all measurements use these."""
import numpy as np
import torch
import torch.nn.functional as F


def sampling_edges(num_atoms):
    """Complete directed ligand graphs in the sampling order of utils/sample_utils.py:40-54 (upper-triangular pairs,
    then the flipped copies); global ligand numbering.  Returns (edge_index [2,E] int64, edge_batch [E] int64)."""
    ei, eb, start = [], [], 0
    for g, n in enumerate(int(v) for v in num_atoms):
        a, b = np.triu_indices(n, k=1)
        half = np.stack([a, b])
        ei.append(np.concatenate([half, half[::-1]], 1) + start)
        eb.append(np.full(n * (n - 1), g))
        start += n
    return (torch.from_numpy(np.concatenate(ei, 1).astype(np.int64)), torch.from_numpy(np.concatenate(eb).astype(np.int64)))


def training_edges(num_atoms):
    """Same graphs in the dst-major order of datasets/transform.py:488-501."""
    ei, eb, start = [], [], 0
    for g, n in enumerate(int(v) for v in num_atoms):
        dst, src = np.repeat(np.arange(n), n), np.tile(np.arange(n), n)
        m = dst != src
        ei.append(np.stack([src[m], dst[m]]) + start)
        eb.append(np.full(int(m.sum()), g))
        start += n
    return (torch.from_numpy(np.concatenate(ei, 1).astype(np.int64)), torch.from_numpy(np.concatenate(eb).astype(np.int64)))


def synthetic_phore(rng, p, n_ex=0):
    """18-dim pharmacophore features as datasets/get_phore_data.py:55-70 builds them:
    13-way type one-hot, alpha, has_norm one-hot(2), EX one-hot(2).  Non-EX types drawn from the 12
    non-EX classes; n_ex extra exclusion spheres (type 12) appended."""
    P = p + n_ex
    types = np.r_[rng.integers(0, 12, size=p), np.full(n_ex, 12)]
    x = np.zeros((P, 18), dtype=np.float32)
    x[np.arange(P), types] = 1.0
    x[:, 13] = np.r_[rng.uniform(0.5, 1.5, size=p), np.full(n_ex, 0.837)]
    has_norm = np.r_[rng.integers(0, 2, size=p), np.zeros(n_ex, dtype=np.int64)]
    x[np.arange(P), 14 + has_norm] = 1.0
    is_ex = (types == 12).astype(np.int64)
    x[np.arange(P), 16 + is_ex] = 1.0
    pos = rng.normal(0.0, 4.0, size=(P, 3)).astype(np.float32)
    pos -= pos.mean(0, keepdims=True)
    nrm = rng.normal(size=(P, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm *= has_norm[:, None].astype(np.float32)
    return x, pos, nrm.astype(np.float32)


def synthetic_batch(seed, n_graphs, n_atoms=30, p_choices=(6, 7, 8), n_ex=0, edge_order="sampling", pos_scale=1.0):
    """Seeded synthetic ligand+pharmacophore batch in the reference's tensor layout (config[1] shapes).
    n_atoms: int or (lo, hi) inclusive range.  Ligand coordinates are N(0, 1) * pos_scale (the reference's initial state is
    N(0, 1); large molecules are spread by pos_scale > 1 in the parity fixtures so that 80 atoms are not packed into a
    unit ball, where near-ties of the k=32 neighbour selection are unavoidable)."""
    rng = np.random.default_rng(seed)
    if isinstance(n_atoms, int):
        na = np.full(n_graphs, n_atoms)
    else:
        na = rng.integers(n_atoms[0], n_atoms[1] + 1, size=n_graphs)
    xs, ps, ns, bs = [], [], [], []
    for g in range(n_graphs):
        p = int(rng.choice(p_choices))
        x, pos, nrm = synthetic_phore(rng, p, n_ex)
        xs.append(x); ps.append(pos); ns.append(nrm); bs.append(np.full(x.shape[0], g))
    phore = dict(x=torch.from_numpy(np.concatenate(xs)), pos=torch.from_numpy(np.concatenate(ps)),
                 norm=torch.from_numpy(np.concatenate(ns)), batch=torch.from_numpy(np.concatenate(bs).astype(np.int64)))
    Nl = int(na.sum())
    batch_node = torch.from_numpy(np.repeat(np.arange(n_graphs), na).astype(np.int64))
    ei, eb = sampling_edges(na) if edge_order == "sampling" else training_edges(na)
    node_cls = torch.from_numpy(rng.integers(0, 12, size=Nl).astype(np.int64))
    edge_cls = torch.from_numpy(rng.integers(0, 6, size=ei.shape[1]).astype(np.int64))
    pos = torch.from_numpy(rng.normal(0.0, 1.0, size=(Nl, 3)).astype(np.float32))
    if pos_scale != 1.0:
        pos = pos * pos_scale
    return dict(num_atoms=torch.from_numpy(na.astype(np.int64)), batch_node=batch_node, edge_index=ei,
                batch_edge=eb, h_node=F.one_hot(node_cls, 12).float(), h_edge=F.one_hot(edge_cls, 6).float(),
                pos=pos, phore=phore, n_graphs=n_graphs)
