"""CPU tests of the training tier (phoregen_b200/training.py): loss and gradients of `compute_loss` against the unmodified
reference's autograd (models/diffusion.py:249-352) for all trainable parameters, and the data-parallel gradient reducer
under gloo with world size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build_model, load_golden
from oracle import phoregen_oracle as O
from phoregen_b200 import training
from phoregen_b200.testing import grad_digest, training_batch_from_synthetic

HAVE_REF = os.path.isdir("/root/reference/models")


class OracleProvider:
    """Index artefacts from the CPU oracle (the production provider is a BatchPlan: CUDA graph kernels)."""

    def __init__(self, data):
        lig, ll, ph = data["ligand"], data["ligand", "ligand"], data["phore"]
        _, self.batch_ctx, self.mask, _, l_idx = O.compose_context(ph.batch, lig.batch)
        self.bond_ctx = l_idx[ll.f_edge_index]

    def triplets(self):
        return O.triplets(self.bond_ctx, self.batch_ctx.numel())

    def knn_graph(self, x, mode):
        return O.knn_graph(x, 32, self.batch_ctx) if mode == 0 else O.knn_graph(x[self.mask], 3, self.batch_ctx[self.mask])


def loss_and_grads(model, data, seed, provider=None):
    model.train()
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(seed)
    loss, terms = model.compute_loss(data, provider=provider)
    loss.backward()
    return loss.detach(), terms, {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad and p.grad is not None}


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_loss_and_gradients_match_unmodified_reference_autograd():
    import yaml
    from oracle.shims.install import EasyDict, install
    install()
    from models.diffusion import PhoreDiff
    mirror, sd = build_model()
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2
    ref = PhoreDiff(cfg.model, "zinc_300")
    ref.load_state_dict(sd, strict=True)
    data = training_batch_from_synthetic(O.synthetic_batch(91, 3, n_atoms=(5, 9), edge_order="training"))
    ref.train()
    torch.manual_seed(3)
    want_loss, want_terms = ref.compute_loss(data)
    want_loss.backward()
    want = {k: p.grad for k, p in ref.named_parameters() if p.requires_grad and p.grad is not None}
    got_loss, got_terms, got = loss_and_grads(mirror, data, 3, provider=OracleProvider(data))
    assert float(got_loss) == pytest.approx(float(want_loss), rel=1e-5)
    for k in want_terms:
        assert got_terms[k] == pytest.approx(want_terms[k], rel=1e-4, abs=1e-6), k
    assert set(got) == set(want) and len(got) > 300
    n_param = sum(v.numel() for v in got.values())
    assert n_param == 5201785                                    # every trainable parameter receives a gradient
    # The key MLPs' output biases (*k_func.net.3.bias) shift every logit of a softmax segment by the same q.b: their true
    # gradient is exactly zero and both sides hold rounding noise (1e-9 of the other gradients) - compared absolutely.
    scale = max(float(v.norm()) for v in want.values())
    worst = 0.0
    for k in want:
        diff, ref_n = float((got[k] - want[k]).norm()), float(want[k].norm())
        if ref_n < 1e-6 * scale:
            assert float(got[k].norm()) < 1e-6 * scale, (k, ref_n, float(got[k].norm()))      # (also: heads gated off by relu in this batch)
            continue
        worst = max(worst, diff / ref_n)
        assert diff / ref_n < 1e-3, (k, diff / ref_n)
    assert 0 < worst < 1e-3


def test_gradient_fixture_from_reference_matches(tmp_path):
    """The committed fingerprint of the reference's gradients (oracle/make_golden.py train) pins this path where the
    reference itself is not available (the GPU box); here on the CPU with the oracle's index artefacts."""
    f = load_golden("train_grads.pt")
    mirror, _ = build_model()
    data = training_batch_from_synthetic(O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=tuple(f["n_atoms"]), edge_order="training"))
    loss, _, got = loss_and_grads(mirror, data, f["torch_seed"], provider=OracleProvider(data))
    assert float(loss) == pytest.approx(f["loss"], rel=1e-5)
    dig = grad_digest(got)
    assert set(dig) == set(f["digest"])
    check_digest(dig, f["digest"])


def check_digest(got, want, rel=1e-3, prj_slack=3.0):
    """Norms within `rel`; projections within `rel` x norm x sqrt-ish slack (a projection is a sum of signed terms); tensors
    whose reference gradient is numerically zero (key output biases: softmax-invariant) must be numerically zero too.
    `prj_slack`: the GPU run accumulates the scatter gradients with atomics in a run-dependent order; the projection of
    a small, cancellation-dominated gradient (e.g. a key-MLP bias, |g| ~ 1e-3 of the largest) moves by a few 1e-3 |g|
    between runs there, so that caller passes a wider slack on the projection (the norm bar stays at `rel`)."""
    scale = max(n for n, _ in want.values())
    for k, (nrm, prj) in want.items():
        if nrm < 1e-6 * scale:
            assert got[k][0] < 1e-5 * scale, k
            continue
        assert got[k][0] == pytest.approx(nrm, rel=rel), k
        assert abs(got[k][1] - prj) <= prj_slack * rel * nrm * np.sqrt(2.0) + 1e-12, k


# ---------------------------------------------------------------- gradient reducer, gloo world size 2
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _reducer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(40, 300), torch.nn.ReLU(), torch.nn.Linear(300, 300), torch.nn.ReLU(), torch.nn.Linear(300, 5))
    torch.manual_seed(0)
    net = mk()
    plain = mk()                                                  # same weights, no reducer: the rank's own gradients
    plain.load_state_dict(net.state_dict())
    unused = torch.nn.Parameter(torch.ones(7))                    # a parameter that never gets a gradient (find_unused_parameters)
    red = training.GradientReducer(list(net.parameters()) + [unused], bucket_mb=0.005)
    assert len(red.buckets) >= 3
    outs = []
    for step in range(2):
        red.zero_grad()
        plain.zero_grad()
        x = torch.randn(16, 40, generator=torch.Generator().manual_seed(100 * step + rank))
        net(x).pow(2).mean().backward()                           # buckets are all-reduced while this runs
        plain(x).pow(2).mean().backward()
        local = torch.cat([p.grad.reshape(-1).clone() for p in plain.parameters()])
        red.finish()
        outs.append((local, torch.cat([p.grad.reshape(-1).clone() for p in net.parameters()]), unused.grad.clone()))
    torch.save(outs, os.path.join(out_dir, f"red{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_reducer_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "red0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "red1.pt", weights_only=False)
    for (l0, a0, u0), (l1, a1, u1) in zip(r0, r1):
        assert not torch.equal(l0, l1)                           # different data per rank
        assert torch.allclose(a0, (l0 + l1) / 2, rtol=1e-5, atol=1e-7) and torch.equal(a0, a1)
        assert torch.equal(u0, torch.zeros(7)) and torch.equal(u1, torch.zeros(7))
