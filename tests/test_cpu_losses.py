"""CPU tests of the forward-only training objective (phoregen_b200/losses.py) against the unmodified reference
`PhoreDiff.compute_loss` (models/diffusion.py:249-352): same random draws under equal seeds, same loss terms for the same
predictions."""
import os

import pytest
import torch

from helpers import build_model
from oracle import phoregen_oracle as O
from phoregen_b200 import losses
from phoregen_b200.testing import training_batch_from_synthetic

HAVE_REF = os.path.isdir("/root/reference/models")
pytestmark = [pytest.mark.reference, pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")]


def _reference_model(sd):
    import yaml
    from oracle.shims.install import EasyDict, install
    install()
    from models.diffusion import PhoreDiff
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2
    ref = PhoreDiff(cfg.model, "zinc_300").eval()
    ref.load_state_dict(sd, strict=True)
    return ref


@pytest.mark.parametrize("seed", [0, 4])
def test_noise_draws_and_loss_terms_match_the_reference(seed):
    mirror, sd = build_model()                       # CPU module: tables and heads only, the forward is injected below
    ref = _reference_model(sd)
    data = training_batch_from_synthetic(O.synthetic_batch(70 + seed, 5, n_atoms=(4, 9), edge_order="training"))
    captured = {}

    def fake_forward(**kw):                          # deterministic stand-in predictions, shared by both sides
        g = torch.Generator().manual_seed(123)
        N, E, G = kw["pos_pert"].shape[0], kw["h_edge_pert"].shape[0], kw["time_step"].numel()
        captured.setdefault("inputs", []).append({k: v.clone() for k, v in kw.items()})
        lo = torch.rand(G, 1, generator=g) * 0.4
        return (torch.randn(N, 12, generator=g), kw["pos_pert"] + 0.1 * torch.randn(N, 3, generator=g),
                torch.randn(E, 6, generator=g), (lo, lo + torch.rand(G, 1, generator=g) * 0.5))

    ref.forward = fake_forward
    torch.manual_seed(seed)
    with torch.no_grad():
        want_total, want = ref.compute_loss(data)
    torch.manual_seed(seed)
    got_total, got = losses.compute_loss(mirror, data, forward=fake_forward)
    a, b = captured["inputs"]
    for k in a:                                      # identical noise levels and perturbed inputs
        assert torch.equal(a[k], b[k]), k
    assert set(got) == set(want)
    for k in want:
        assert got[k] == pytest.approx(want[k], rel=1e-6, abs=1e-7), k
    assert float(got_total) == pytest.approx(float(want_total), rel=1e-6)


def test_antithetic_time_sampling():
    torch.manual_seed(1)
    t = losses.sample_time(7, 1000, "cpu")
    assert t.shape == (7,) and torch.equal(t[4:], 999 - t[:3])
