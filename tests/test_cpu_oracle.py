"""CPU tests: the oracle restatement against the golden fixtures produced by the unmodified reference
(and against the reference itself where /root/reference is mounted)."""
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close, build_model, load_golden
from oracle import phoregen_oracle as O

HAVE_REF = os.path.isdir("/root/reference/models")


@pytest.fixture(scope="module")
def sd():
    from phoregen_b200.testing import state_dict_digest
    _, sd = build_model()
    assert load_golden("meta.pt")["state_dict_digest"] == state_dict_digest(sd)
    return sd


def test_state_dict_layout_matches_reference_checkpoint_contract(sd):
    meta = load_golden("meta.pt")
    assert len(sd) == 641
    assert sorted((k, tuple(v.shape)) for k, v in sd.items()) == meta["keys"]


def test_oracle_graph_artefacts_match_reference_golden():
    g = load_golden("graph.pt")
    ei, eb = O.make_edge_data(g["num_atoms"])
    assert torch.equal(ei, g["edge_index"]) and torch.equal(eb, g["edge_batch"])
    assert torch.equal(O.knn_graph(g["x"], 32, g["batch"]), g["knn32"])
    m = g["mask_ligand"]
    assert torch.equal(O.knn_graph(g["x"][m], 3, g["batch"][m]), g["knn3"])
    for got, want in zip(O.triplets(g["bond_ctx"], g["x"].shape[0]), g["triplets"]):
        assert torch.equal(got, want)


@pytest.mark.parametrize("name", ["forward_small.pt", "forward_ex.pt"])
def test_oracle_forward_matches_reference_golden(sd, name):
    f = load_golden(name)
    b = O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=f["n_atoms"], n_ex=f["n_ex"])
    ph = b["phore"]
    stages = []
    v, pos, e, cnt = O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"],
                                         b["batch_edge"], torch.tensor(f["times"]), ph["x"], ph["pos"], ph["norm"], ph["batch"],
                                         stages=stages)
    tight = dict(rtol=1e-5, atol=2e-5)
    assert_close(v, f["pred_node"], "logits_node", **tight)
    assert_close(pos, f["pred_pos"], "pos", **tight)
    assert_close(e, f["pred_edge"], "logits_edge", **tight)
    assert_close(cnt[0], f["count_l"], "count_l", **tight)
    if "layer0" in f:
        assert_close(stages[0]["h_phore_emb"], f["h_phore_emb"], "h_phore_emb", **tight)
        for got, want, what in zip((stages[2]["h"], stages[2]["h_bond"], stages[2]["x"]), f["layer0"], ("h", "h_bond", "x")):
            assert_close(got, want, "layer0 " + what, **tight)


def test_oracle_transitions_match_reference_golden(sd):
    f = load_golden("transition.pt")
    for kind in ("node", "edge"):
        d = f[kind]
        post = O.q_v_posterior(sd[f"{kind}_transition.q_mats"], sd[f"{kind}_transition.transpopse_q_onestep_mats"],
                               F.log_softmax(d["pred"], -1), d["log_vt"], f["t"], d["batch"])
        assert torch.equal(post, d["post"])
        assert torch.equal(O.log_sample_categorical(post, d["uniform"]), d["cls"])
    p = f["pos"]
    assert torch.equal(O.pos_prev_from_recon(sd, p["x_t"], p["x_recon"], f["t"], p["batch"], p["normal"], p["grad"]), p["x_prev"])


def test_oracle_reverse_steps_match_reference_golden(sd):
    f = load_golden("reverse_steps.pt")
    b = O.synthetic_batch(f["seed"], 3, n_atoms=(8, 11))
    topo = dict(batch_node=b["batch_node"], edge_index=b["edge_index"], batch_edge=b["batch_edge"], n_graphs=3)
    st = f["init"]
    for step, d, want in zip(f["steps"], f["draws"], f["outs"]):
        st, out = O.reverse_step(sd, st, step, topo, b["phore"], d)
        assert torch.equal(out["node_cls"], want["node_cls"]) and torch.equal(out["edge_cls"], want["edge_cls"])
        assert_close(st["pos"], want["pos"], "pos", rtol=1e-5, atol=1e-5)
        assert_close(st["log_edge"], want["log_edge"], "log_edge", rtol=1e-5, atol=1e-5)


def test_synthetic_batch_is_reproducible():
    a, b = O.synthetic_batch(2032, 3, n_atoms=30), O.synthetic_batch(2032, 3, n_atoms=30)
    assert torch.equal(a["pos"], b["pos"]) and torch.equal(a["phore"]["x"], b["phore"]["x"])
    assert a["edge_index"].shape == (2, 3 * 870) and a["phore"]["x"].shape[1] == 18


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_oracle_matches_unmodified_reference_live(sd):
    import yaml
    from oracle.shims.install import EasyDict, install
    install()
    from models.diffusion import PhoreDiff
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2
    ref = PhoreDiff(cfg.model, "zinc_300").eval()
    # the mirror's schedule tables are bit-identical to the reference's own constructor output
    own = ref.state_dict()
    for k in own:
        if "transition" in k or k.endswith(("offset", "coeff", "freq_bands")):
            assert torch.equal(own[k], sd[k]), k
    ref.load_state_dict(sd, strict=True)
    b = O.synthetic_batch(55, 2, n_atoms=(5, 8), n_ex=10)
    ph = b["phore"]
    t = torch.tensor([321, 0])
    with torch.no_grad():
        want = ref(b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], t, ph["x"], ph["pos"], ph["norm"], ph["batch"])
    got = O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], t, ph["x"], ph["pos"], ph["norm"], ph["batch"])
    for g_, w_ in zip(got[:3], want[:3]):
        assert_close(g_, w_, "forward", rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("case", ["n35", "n48"])
def test_oracle_forward_matches_reference_golden_big_molecules(sd, case):
    """The fixtures that pin the chunked-segment kernels (n - 2 > 32): the oracle agrees with the unmodified reference there too."""
    f = load_golden("forward_big.pt")["cases"][case]
    b = O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=f["n_atoms"], p_choices=f["p_choices"], pos_scale=f["pos_scale"])
    ph = b["phore"]
    v, pos, e, _ = O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"],
                                       torch.tensor(f["times"]), ph["x"], ph["pos"], ph["norm"], ph["batch"])
    tight = dict(rtol=1e-5, atol=2e-5)
    assert_close(v, f["pred_node"], "logits_node", **tight)
    assert_close(pos, f["pred_pos"], "pos", **tight)
    assert_close(e, f["pred_edge"], "logits_edge", **tight)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_oracle_sampling_entry_pieces_match_unmodified_reference_live(sd):
    """D2 / T4 / T5 pieces of PhoreDiff.sample: the atom-count interval of sample_nodes, sample_init under an injected
    uniform draw, and the guidance gradient (autograd of the reference's energies) for a single pharmacophore."""
    import yaml
    from oracle.shims.install import EasyDict, install
    install()
    import models.common as rc
    from models.diffusion import PhoreDiff
    from utils.sample_utils import compute_batch_atom_prox_loss, compute_batch_center_prox_loss, make_edge_data
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2
    ref = PhoreDiff(cfg.model, "zinc_300").eval()
    ref.load_state_dict(sd, strict=True)
    # ---- T4
    g = torch.Generator().manual_seed(4)
    for trans, K in ((ref.node_transition, 12), (ref.edge_transition, 6)):
        u = torch.rand(50, K, generator=g)
        orig = torch.rand_like
        torch.rand_like = lambda x: u.to(x.dtype)
        try:
            cls, onehot, log_vt = trans.sample_init(50)
        finally:
            torch.rand_like = orig
        got_cls, got_log = O.sample_init(trans.init_prob, u)
        assert torch.equal(got_cls, cls) and torch.allclose(got_log, log_vt.float())
    # ---- D2: interval of sample_nodes (diffusion.py:356-380) through the oracle's encoder + count heads
    import numpy as np
    x, pos, _ = O.synthetic_phore(np.random.default_rng(3), 7, 30)
    x, pos = torch.from_numpy(x), torch.from_numpy(pos)
    batch = torch.zeros(x.shape[0], dtype=torch.long)
    cl, cu = O.predict_atom_count(sd, O.phore_encode(sd, x, pos, batch), batch, x, 1)
    lo, hi = int((cl * 74 + 4).round()), int((cu * 74 + 4).round())
    from oracle.shims.install import HeteroData
    data = HeteroData()
    data["phore"].x, data["phore"].pos = x, pos
    seen = []
    import utils.sample_utils as su
    import models.diffusion as md
    orig = md.sample_from_interval
    md.sample_from_interval = lambda l, u_, bs, mode="uniform", scale=4.0: (seen.append((l, u_)), orig(l, u_, bs, mode=mode, scale=scale))[1]
    try:
        with torch.no_grad():
            ref.sample_nodes(data, 8, "cpu")
    finally:
        md.sample_from_interval = orig
    assert seen == [(lo, hi)]
    # ---- T5: closed-form gradient against autograd of the reference's energies (sample_utils.py:135-165)
    na = torch.tensor([5, 7, 4])
    ei, eb = make_edge_data(na)
    bn = torch.repeat_interleave(torch.arange(3), na)
    g = torch.Generator().manual_seed(8)
    xt = (torch.randn(int(na.sum()), 3, generator=g) * 1.5).requires_grad_(True)
    onehot = F.one_hot(torch.randint(0, 6, (ei.shape[1],), generator=g), 6).float()
    centre = torch.randn(3, generator=g)
    e1 = compute_batch_atom_prox_loss(xt, bn, onehot, ei, eb, min_d=1.2, max_d=1.9)
    e2 = compute_batch_center_prox_loss(xt, bn, centre)
    want = torch.autograd.grad(e1, xt)[0] + torch.autograd.grad(e2, xt)[0]
    got = O.guidance_grad(xt.detach(), bn, onehot.argmax(-1), ei, eb, [dict(type="atom_prox", min_d=1.2, max_d=1.9), dict(type="center_prox")], centre, 3)
    assert_close(got, want, "guidance gradient", rtol=1e-5, atol=1e-6)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_reference_side_caller_code_runs_on_the_mirror():
    """Reference-side caller code against the drop-in's module tree: `freeze_parameters` (utils/training_utils.py:18-26)
    walks denoiser.num_layers / base_block[i].pos_layer_with_edge|bond, `get_parameter_number` (:12-15) counts the
    parameters, and the reference's own PhoreDiff constructs with OUR get_denoiser_net / get_phore_encoder swapped in
     and still loads the 641-key checkpoint strictly."""
    import yaml
    from oracle.shims.install import EasyDict, install
    install()
    # utils/training_utils.py imports the whole dataset stack (lmdb, rdkit, PyG transforms) at module level: run the two
    # functions' own source, taken from the reference file at test time, without the module's imports
    import ast
    src = open("/root/reference/utils/training_utils.py").read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("freeze_parameters", "get_parameter_number")]
    ns = {}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "/root/reference/utils/training_utils.py", "exec"), ns)
    freeze_parameters, get_parameter_number = ns["freeze_parameters"], ns["get_parameter_number"]
    mirror, sd = build_model()
    assert "5.5688 M" in get_parameter_number(mirror) and "5.2018 M" in get_parameter_number(mirror)
    frozen = freeze_parameters(mirror, EasyDict(freeze_pos=True))
    pos = [p for l in frozen.denoiser.base_block for m_ in (l.pos_layer_with_edge, l.pos_layer_with_bond) for p in m_.parameters()]
    assert len(pos) == 6 * 2 * 18 and not any(p.requires_grad for p in pos)
    assert all(p.requires_grad for p in frozen.denoiser.base_block[0].bond_layer.parameters())
    # the reference's PhoreDiff with our factories
    import models
    import models.diffusion as md
    from phoregen_b200 import modules as ours
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2
    orig = (md.get_denoiser_net, md.get_phore_encoder)
    md.get_denoiser_net, md.get_phore_encoder = ours.get_denoiser_net, ours.get_phore_encoder
    try:
        hybrid = md.PhoreDiff(cfg.model, "zinc_300")
    finally:
        md.get_denoiser_net, md.get_phore_encoder = orig
    assert isinstance(hybrid.denoiser, ours.UniTransformerO2TwoUpdateGeneralBond)
    hybrid.load_state_dict(sd, strict=True)
    assert len(hybrid.state_dict()) == 641
