"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol the header declares;
the weight packer and the library agree on the slot table; host-side helpers."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT, build_model


def test_library_exports_every_declared_symbol():
    from phoregen_b200 import _lib
    header = open(os.path.join(ROOT, "include", "phoregen_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib.pg_version() >= 100


def test_weight_packer_fills_every_slot():
    from phoregen_b200 import _lib
    from phoregen_b200.weights import build_blob, pack_state_dict
    _, sd = build_model()
    table = _lib.slot_table()
    assert len(table) == len({n for n, _ in table})
    blob, off = build_blob(sd)
    assert blob.dtype == np.float32 and np.all(off % 4 == 0) and np.isfinite(blob).all()
    packed = pack_state_dict(sd)
    # the fold of dire_embedding and the type-selected smearing slices reproduce the reference's first Linear exactly
    rng = np.random.default_rng(0)
    W1 = sd["denoiser.base_block.2.node_layer_with_edge.hk_func.net.0.weight"].double().numpy()
    b1 = sd["denoiser.base_block.2.node_layer_with_edge.hk_func.net.0.bias"].double().numpy()
    Wde = sd["denoiser.base_block.2.dire_embedding.weight"].double().numpy()
    bde = sd["denoiser.base_block.2.dire_embedding.bias"].double().numpy()
    smear, dots, hd, hs = rng.random(20), rng.normal(size=3), rng.normal(size=128), rng.normal(size=128)
    for t in range(4):
        onehot = np.eye(4)[t]
        edge_feat = np.concatenate([np.outer(onehot, smear).reshape(-1), onehot, Wde @ dots + bde])
        want = W1 @ np.concatenate([edge_feat, hd, hs]) + b1
        tab = packed["L2.nk.tab_k"][t]
        n1, nb = packed["L2.n1.wt"], packed["L2.n1.b"]
        got = (smear @ tab[:20] + tab[20] + dots @ tab[21:24] + hd @ n1[:, 0:128] + nb[0:128] + hs @ n1[:, 128:256] + nb[128:256])
        # ... up to the channel mean, which the packer removes from every first-Linear block (the LayerNorm that follows
        # subtracts it anyway; weights._center_first_linears)
        np.testing.assert_allclose(got, want - want.mean(), rtol=1e-12, atol=1e-12)
        assert abs(got.mean()) < 1e-13


def _bf16_image_to_f64(carrier, n_out, k=128):
    """[hi|lo][n_out][k] bf16 carried in fp32 words (weights.bf16_split) -> hi + lo as float64 [n_out][k]."""
    bits = np.asarray(carrier, dtype=np.float32).view(np.uint16).reshape(2, n_out, k)
    f = (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    return f[0] + f[1]


def test_layernorm_gain_fold_of_the_tensor_core_weights():
    """weights.build_blob folds positive LayerNorm gains into the bf16 W2 images (relu(g*x + b) = g*relu(x + b/g)) and falls
    back to the plain image + beta when a gain is not positive; either way image . relu(.) reproduces W2 . relu(g*x + b)."""
    from phoregen_b200 import _lib
    from phoregen_b200.weights import build_blob
    _, sd = build_model()
    key = "denoiser.base_block.1.bond_ffn.hv_func.net.1.weight"
    key = key if key in sd else next(k for k in sorted(sd) if k.endswith("hv_func.net.1.weight") and ".base_block.1." in k)
    sd2 = dict(sd)
    g = sd[key].clone(); g[5] = -g[5]
    sd2[key] = g
    rng = np.random.default_rng(1)
    xn = rng.normal(size=128)
    for state, expect_all_folded in ((sd, True), (sd2, False)):
        blob, off = build_blob(state)
        slot = {n: (int(o), int(m)) for (n, m), o in zip(_lib.slot_table(), off)}
        get = lambda n: blob[slot[n][0]: slot[n][0] + slot[n][1]]
        folded = []
        for S in ("L1.tr.", "L1.nb.", "L1.pb.", "L1.nk.", "L1.pk."):
            for i, kv in enumerate("kv"):
                n_out = slot[S + f"w2{kv}"][1] // 128
                w2 = get(S + f"w2{kv}").reshape(n_out, 128).astype(np.float64)
                gam, bet = get(S + f"ln{kv}_g").astype(np.float64), get(S + f"ln{kv}_b").astype(np.float64)
                fold = get(S + "fold")[i] > 0.5
                folded.append(bool(fold))
                assert fold == bool(np.all(gam > 1e-3))
                img = _bf16_image_to_f64(get(S + f"w2{kv}.bf"), n_out)
                bf = get(S + f"ln{kv}_bf").astype(np.float64)
                hid = np.maximum(xn + bf, 0.0) if fold else np.maximum(gam * xn + bf, 0.0)
                want = w2 @ np.maximum(gam * xn + bet, 0.0)
                np.testing.assert_allclose(img @ hid, want, rtol=0, atol=3e-5 * (1.0 + np.abs(want).max()))
        assert all(folded) == expect_all_folded


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from phoregen_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PhoreGenLibraryError, match="no CPU / PyTorch fallback"):
        _lib._load()


def test_cpu_device_is_refused():
    from phoregen_b200 import _lib
    from phoregen_b200.engine import PackedModel
    _, sd = build_model()
    with pytest.raises(_lib.PhoreGenLibraryError, match="CUDA devices only"):
        PackedModel(sd, "cpu")


def test_topology_from_context():
    from oracle import phoregen_oracle as O
    from phoregen_b200.modules import topology_from_context
    bp = torch.tensor([0, 0, 0, 1, 1, 2])
    bl = torch.tensor([0, 0, 1, 1, 1, 1, 2, 2])
    _, batch, mask, _, _ = O.compose_context(bp, bl)
    p, n = topology_from_context(batch, mask)
    assert p.tolist() == [3, 2, 1] and n.tolist() == [2, 4, 2]
    with pytest.raises(ValueError):
        topology_from_context(batch, ~mask)


def test_schedules_shapes_and_invariants():
    from phoregen_b200 import schedules
    b = schedules.beta_schedule("segment", 1000, time_segment=[600, 400],
                                segment_diff=[dict(scale_start=0.9999, scale_end=0.001, width=3), dict(scale_start=0.001, scale_end=0.0001, width=2)])
    assert b.shape == (1000,) and (b >= 0).all() and (b <= 1).all()
    tabs, prior = schedules.categorical_tables(b, 6, "absorb")
    np.testing.assert_allclose(tabs["q_mats"].sum(-1), 1.0, atol=1e-9)       # rows of a transition matrix
    np.testing.assert_allclose(prior.sum(), 1.0)
    g = schedules.gaussian_tables(schedules.beta_schedule("advance", 1000, scale_start=0.9999, scale_end=0.0001, width=3))
    assert g["std"][0] == 0.0 and np.all(np.diff(g["alphas_bar"]) <= 0)


def test_mma_issue_stays_on_the_uniform_datapath():
    """The tcgen05 kernels issue their MMAs warp-collectively (pg_tc.cuh umma_*_w): descriptor arithmetic on the uniform
    datapath, one elected lane issues.  Issued from `if (lane == 0)` inside a role branch on `threadIdx.x >> 5` instead,
    ptxas feeds every UTCHMMA through an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~110 cycles per MMA; the
    triplet kernel was 25 % slower).  The SASS of the built library must not contain that pattern."""
    import shutil
    import subprocess
    from phoregen_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) and not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            continue
        if cur and any(k in cur for k in ("trip_tc_kernel", "bond_tc_kernel", "knn_tc_kernel", "gemm_tc_kernel")):
            c = counts.setdefault(cur, [0, 0])
            c[0] += "UTCHMMA" in line
            c[1] += "R2UR.BROADCAST" in line
    assert counts, "no tensor-core kernels found in the library"
    assert all(c[0] > 0 for c in counts.values())                      # tcgen05.mma present in every one of them
    assert sum(c[1] for c in counts.values()) == 0, {k: v for k, v in counts.items() if v[1]}


def test_triplet_kernel_keeps_its_register_budget():
    """trip_tc_kernel runs 640 threads per SM at 96 registers (row warps 104 after setmaxnreg); every instantiation must fit
    without spilling (a spilled value turns an asynchronous TMEM / global load into a blocking one in the row warps' loop).
    ptxas' stack frame of these kernels is 16 bytes with no spill; a spill shows up as a larger frame."""
    import re
    import shutil
    import subprocess
    from phoregen_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) and not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout.splitlines()
    seen = 0
    for i, line in enumerate(out):
        if "Function" in line and "trip_tc_kernel" in line:
            m = re.search(r"REG:(\d+) STACK:(\d+)", out[i + 1])
            assert m, out[i + 1]
            assert int(m.group(1)) <= 96 and int(m.group(2)) <= 16, (line.strip(), out[i + 1].strip())
            seen += 1
    assert seen == 6          # {single-chunk, chunked} x {bf16x3, fp16, fp16 hi/lo} key paths
