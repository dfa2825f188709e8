"""Shared test helpers: model construction with the reproducible weights, golden loading, tolerance."""
import os

import torch

from phoregen_b200.testing import MODEL_CONFIG, random_state_dict, state_dict_digest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
# BASELINE.json north_star: per-step outputs within 1e-3 relative / 1e-4 absolute of the fp32 reference forward
RTOL, ATOL = 1e-3, 1e-4


def load_golden(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def build_model(device=None, seed=0):
    from phoregen_b200.diffusion import PhoreDiff
    m = PhoreDiff(MODEL_CONFIG, "zinc_300")
    sd = random_state_dict(m, seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    if device is not None:
        m = m.to(device)
    return m, sd


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol={rtol} atol={atol}; "
                                 f"max abs err {err.max().item():.3e}, max ref {b.abs().max().item():.3e}")


def report(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    return f"max_abs={err.max().item():.3e} max_rel_to_tol={(err / (ATOL + RTOL * b.abs())).max().item():.3f} ref_max={b.abs().max().item():.3e}"
