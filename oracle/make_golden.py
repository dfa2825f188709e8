"""TEST INFRASTRUCTURE — generate tests/golden/*.pt by running the UNMODIFIED reference (imported from
/root/reference through oracle/shims) on seeded synthetic inputs.  Run here (the GPU box has no /root/reference):

    python -m oracle.make_golden

Weights come from phoregen_b200.testing.random_state_dict (numpy PCG64, reproducible anywhere) loaded into the
reference model with strict=True; the fixture stores the state_dict digest so a consumer can prove it rebuilt the
same weights.  Inputs are regenerated from seeds by oracle.phoregen_oracle.synthetic_batch.
"""
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.shims.install import EasyDict, install  # noqa: E402

install()
from oracle import phoregen_oracle as O  # noqa: E402
from phoregen_b200.diffusion import PhoreDiff as MirrorPhoreDiff  # noqa: E402
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict, state_dict_digest  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def reference_model(seed=0):
    import models.common as rc
    from models.diffusion import PhoreDiff
    cfg = EasyDict(yaml.safe_load(open("/root/reference/configs/train_lig-phore.yml")))
    cfg.model.phore_feat_dim += 2                      # sample_all.py:41-43
    ref = PhoreDiff(cfg.model, "zinc_300").eval()
    sd = random_state_dict(MirrorPhoreDiff(MODEL_CONFIG, "zinc_300"), seed)
    ref.load_state_dict(sd, strict=True)
    return ref, sd, rc


MIN_MARGIN = 1e-3     # smallest accepted kNN selection margin (squared Angstrom) of a fixture, see oracle.knn_margin


def well_conditioned_seed(sd, seed, n_graphs, n_atoms, times, n_ex=0):
    """First seed >= `seed` whose forward pass never comes close to a kNN tie (the graph is discontinuous there)."""
    while True:
        b = O.synthetic_batch(seed, n_graphs, n_atoms=n_atoms, n_ex=n_ex)
        ph = b["phore"]
        stages = []
        O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"],
                            torch.tensor(times), ph["x"], ph["pos"], ph["norm"], ph["batch"], stages=stages)
        m = O.forward_knn_margin(stages)
        if m >= MIN_MARGIN:
            return seed, m
        print(f"  seed {seed}: kNN margin {m:.2e} too small, trying the next one")
        seed += 1


def big_forward_fixture(ref, sd, seed, n_graphs, n_atoms, times, p_choices=(10, 11, 12), pos_scale=2.5, min_margin=5e-4):
    """Forward outputs of the unmodified reference for molecules beyond one 32-row attention chunk (n up to 80)."""
    while True:
        b = O.synthetic_batch(seed, n_graphs, n_atoms=n_atoms, p_choices=p_choices, pos_scale=pos_scale)
        ph = b["phore"]
        stages = []
        O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"],
                            torch.tensor(times), ph["x"], ph["pos"], ph["norm"], ph["batch"], stages=stages)
        margin = O.forward_knn_margin(stages)
        if margin >= min_margin:
            break
        print(f"  big seed {seed}: kNN margin {margin:.2e} too small, trying the next one", flush=True)
        seed += 1
    with torch.no_grad():
        out = ref(b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"],
                  torch.tensor(times, dtype=torch.long), ph["x"], ph["pos"], ph["norm"], ph["batch"])
    return dict(seed=seed, knn_margin=margin, n_graphs=n_graphs, n_atoms=n_atoms, p_choices=p_choices, pos_scale=pos_scale,
                times=times, pred_node=out[0], pred_pos=out[1], pred_edge=out[2])


def big_main():
    """python -m oracle.make_golden big  ->  tests/golden/forward_big.pt (minutes of CPU time: the reference evaluates
    437-wide MLPs on 493k triplets per layer at n = 80)."""
    torch.set_num_threads(os.cpu_count())
    ref, sd, _ = reference_model(0)
    cases = {}
    for name, seed, G, na, times in (("n35", 201, 1, 35, [412]), ("n48", 202, 1, 48, [37]), ("n64", 203, 1, 64, [871]),
                                     ("n78", 204, 1, 78, [5]), ("n80", 205, 1, 80, [640]),
                                     ("ragged20_40", 206, 8, (20, 40), [999, 700, 512, 300, 123, 45, 1, 0])):
        cases[name] = big_forward_fixture(ref, sd, seed, G, na, times, p_choices=(10, 11, 12) if G == 1 else (6, 7, 8))
        print(name, "seed", cases[name]["seed"], "margin", cases[name]["knn_margin"], flush=True)
        torch.save(dict(state_dict_digest=state_dict_digest(sd), cases=cases), os.path.join(GOLD, "forward_big.pt"))


def train_main():
    """python -m oracle.make_golden train -> tests/golden/train_grads.pt: loss and a per-tensor fingerprint of the gradients of
    the unmodified reference's compute_loss (models/diffusion.py:249-352) + autograd, for the training-tier parity tests."""
    from phoregen_b200.testing import grad_digest, training_batch_from_synthetic
    ref, sd, _ = reference_model(0)
    ref.train()
    seed, n_graphs, n_atoms, torch_seed = 92, 4, (6, 12), 5
    data = training_batch_from_synthetic(O.synthetic_batch(seed, n_graphs, n_atoms=n_atoms, edge_order="training"))
    torch.manual_seed(torch_seed)
    loss, terms = ref.compute_loss(data)
    loss.backward()
    grads = {k: p.grad for k, p in ref.named_parameters() if p.requires_grad and p.grad is not None}
    fix = dict(seed=seed, n_graphs=n_graphs, n_atoms=list(n_atoms), torch_seed=torch_seed, loss=float(loss), terms=terms,
               digest=grad_digest(grads), n_params=sum(g.numel() for g in grads.values()), state_dict_digest=state_dict_digest(sd))
    torch.save(fix, os.path.join(GOLD, "train_grads.pt"))
    print("train_grads.pt: loss", float(loss), "tensors", len(grads), "params", fix["n_params"])


def phores_main():
    """python -m oracle.make_golden phores -> tests/golden/forward_phores.pt: forward outputs of the unmodified reference on each
    of its 10 shipped sampling pharmacophores (data/phores_for_sampling, 44-99 nodes of which 40-94 exclusion spheres; read by
    the reference's own PhoreData_New), two 14- / 17-atom ligands per pharmacophore, plus the sample_nodes interval."""
    import json
    from datasets.get_phore_data import PhoreData_New
    ref, sd, _ = reference_model(0)
    index = json.load(open("/root/reference/data/phores_for_sampling/file_index.json"))
    files = [os.path.join("/root/reference", p.lstrip("./")) for p in index]
    ds = PhoreData_New(files, center="phore", data_name="zinc_300")
    cases = {}
    for i, path in enumerate(files):
        d = ds.get(i)
        px, ppos, pnorm = d["phore"].x, d["phore"].pos, d["phore"].norm
        P = px.shape[0]
        seed = 400 + i
        while True:
            b = O.synthetic_batch(seed, 2, n_atoms=(14, 17), pos_scale=2.0)
            ph = dict(x=px.repeat(2, 1), pos=ppos.repeat(2, 1), norm=pnorm.repeat(2, 1), batch=torch.repeat_interleave(torch.arange(2), P))
            t = torch.tensor([650, 40])
            stages = []
            O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], t,
                                ph["x"], ph["pos"], ph["norm"], ph["batch"], stages=stages)
            margin = O.forward_knn_margin(stages)
            if margin >= 5e-4:
                break
            print(f"  {os.path.basename(path)} seed {seed}: kNN margin {margin:.2e} too small, next", flush=True)
            seed += 50
        with torch.no_grad():
            out = ref(b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], t, ph["x"], ph["pos"], ph["norm"], ph["batch"])
            seen = []
            import models.diffusion as md
            orig = md.sample_from_interval
            md.sample_from_interval = lambda l, u_, bs, mode="uniform", scale=4.0: (seen.append((l, u_)), orig(l, u_, bs, mode=mode, scale=scale))[1]
            try:
                ref.sample_nodes(d, 4, "cpu")
            finally:
                md.sample_from_interval = orig
        cases[os.path.basename(path)] = dict(seed=seed, knn_margin=margin, n_phore=P, times=[650, 40], pred_node=out[0], pred_pos=out[1], pred_edge=out[2],
                                             count_l=out[3][0], count_u=out[3][1], interval=seen[0])
        print(os.path.basename(path), "nodes", P, "seed", seed, "margin", margin, "interval", seen[0], flush=True)
    torch.save(dict(state_dict_digest=state_dict_digest(sd), cases=cases), os.path.join(GOLD, "forward_phores.pt"))


def forward_fixture(ref, seed, n_graphs, n_atoms, times, n_ex=0, stages=True, sd=None):
    seed, margin = well_conditioned_seed(sd, seed, n_graphs, n_atoms, times, n_ex)
    b = O.synthetic_batch(seed, n_graphs, n_atoms=n_atoms, n_ex=n_ex)
    ph = b["phore"]
    t = torch.tensor(times, dtype=torch.long)
    rec = {}
    hooks = []
    if stages:
        def mk(l):
            def hook(mod, inp, out):
                rec[f"layer{l}"] = [o.detach().clone() for o in out]     # (h, h_bond, x)
            return hook
        for l in (0, 5):
            hooks.append(ref.denoiser.base_block[l].register_forward_hook(mk(l)))
        hooks.append(ref.phore_encoder.register_forward_hook(lambda m, i, o: rec.__setitem__("h_phore_emb", o.detach().clone())))
    with torch.no_grad():
        out = ref(b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], t,
                  ph["x"], ph["pos"], ph["norm"], ph["batch"])
    for h in hooks:
        h.remove()
    fix = dict(seed=seed, knn_margin=margin, n_graphs=n_graphs, n_atoms=n_atoms, n_ex=n_ex, times=times, pred_node=out[0], pred_pos=out[1],
               pred_edge=out[2], count_l=out[3][0], count_u=out[3][1])
    fix.update(rec)
    return fix


def transition_fixture(ref, rc, seed):
    g = torch.Generator().manual_seed(seed)
    G = 5
    t = torch.tensor([999, 640, 333, 1, 0])
    out = {"t": t}
    for kind, K, rows_per in (("node", 12, 7), ("edge", 6, 11)):
        trans = getattr(ref, f"{kind}_transition")
        batch = torch.repeat_interleave(torch.arange(G), rows_per)
        pred = torch.randn(batch.numel(), K, generator=g) * 2.0
        log_vt = torch.log_softmax(torch.randn(batch.numel(), K, generator=g) * 3.0, -1)
        u = torch.rand(batch.numel(), K, generator=g)
        post = trans.q_v_posterior(torch.log_softmax(pred, -1), log_vt, t, batch, v0_prob=True)
        orig = torch.rand_like
        torch.rand_like = lambda x: u                   # inject the draw into common.py:426
        try:
            cls = rc.log_sample_categorical(post)
        finally:
            torch.rand_like = orig
        # margin of the arg-max (fixtures with near ties are useless for a bit-exactness check)
        gum = -torch.log(-torch.log(u + 1e-30) + 1e-30) + post
        top2 = gum.topk(2, -1).values
        out[kind] = dict(batch=batch, pred=pred, log_vt=log_vt, uniform=u, post=post, cls=cls,
                         margin=(top2[:, 0] - top2[:, 1]))
    # position posterior (transition.py:44-63)
    batch = torch.repeat_interleave(torch.arange(G), 6)
    x_t, x0 = torch.randn(batch.numel(), 3, generator=g), torch.randn(batch.numel(), 3, generator=g)
    z = torch.randn(batch.numel(), 3, generator=g)
    grad = 0.01 * torch.randn(batch.numel(), 3, generator=g)
    orig = torch.randn_like
    torch.randn_like = lambda x: z
    try:
        xp = ref.pos_transition.get_prev_from_recon(x_t=x_t, x_recon=x0, t=t, batch=batch, energy_grad=grad)
    finally:
        torch.randn_like = orig
    out["pos"] = dict(batch=batch, x_t=x_t, x_recon=x0, normal=z, grad=grad, x_prev=xp)
    return out


def graph_fixture():
    from models.uni_denoiser import BondUpdateLayer
    from torch_geometric.nn import knn_graph
    from utils.sample_utils import make_edge_data
    g = torch.Generator().manual_seed(11)
    na = torch.tensor([5, 9, 4, 12])
    npn = torch.tensor([6, 40, 3, 7])
    ei, eb = make_edge_data(na)
    # context coordinates with exact duplicates (HD+HA on one atom are common in real .phore files)
    sizes = (na + npn).tolist()
    batch = torch.repeat_interleave(torch.arange(4), na + npn)
    x = torch.randn(int(sum(sizes)), 3, generator=g) * 3.0
    x[3] = x[1]; x[20] = x[30]; x[21] = x[30]
    x[60:64] = torch.round(x[60:64])                   # lattice points -> equidistant ties
    knn32 = knn_graph(x, k=32, batch=batch, flow="source_to_target")
    mask = torch.cat([torch.cat([torch.zeros(p, dtype=torch.bool), torch.ones(n, dtype=torch.bool)]) for n, p in zip(na.tolist(), npn.tolist())])
    knn3 = knn_graph(x[mask], k=3, batch=batch[mask])
    # triplets on the context-numbered bond index
    lig_rows = mask.nonzero()[:, 0]
    bond_ctx = lig_rows[ei]
    trip = BondUpdateLayer.triplets(bond_ctx, x.shape[0])[2:]
    return dict(num_atoms=na, num_phore=npn, edge_index=ei, edge_batch=eb, x=x, batch=batch, mask_ligand=mask,
                knn32=knn32, knn3=knn3, bond_ctx=bond_ctx, triplets=[t.clone() for t in trip])


def reverse_steps_fixture(ref, rc, seed, steps=(999, 998, 997), sd=None):
    """Loop body of models/diffusion.py:432-517 driven through the reference's own methods with injected draws."""
    while True:
        fix = _reverse_steps_fixture(ref, rc, seed, steps, sd)
        if fix["knn_margin"] >= MIN_MARGIN:
            return fix
        print(f"  reverse steps seed {seed}: kNN margin {fix['knn_margin']:.2e} too small, trying the next one")
        seed += 1


def _reverse_steps_fixture(ref, rc, seed, steps, sd):
    b = O.synthetic_batch(seed, 3, n_atoms=(8, 11))
    ph = b["phore"]
    g = torch.Generator().manual_seed(seed)
    Nl, Eb = b["h_node"].shape[0], b["h_edge"].shape[0]
    state = dict(h_node=b["h_node"], pos=b["pos"], h_edge=b["h_edge"],
                 log_node=torch.log(b["h_node"].clamp(min=1e-30)), log_edge=torch.log(b["h_edge"].clamp(min=1e-30)))
    init = {k: v.clone() for k, v in state.items()}
    draws, outs = [], []
    margin = float("inf")
    for step in steps:
        t = torch.full((3,), step, dtype=torch.long)
        stg = []
        O.phorediff_forward(sd, state["h_node"], state["pos"], b["batch_node"], state["h_edge"], b["edge_index"], b["batch_edge"], t,
                            ph["x"], ph["pos"], ph["norm"], ph["batch"], stages=stg)
        margin = min(margin, O.forward_knn_margin(stg))
        d = dict(u_node=torch.rand(Nl, 12, generator=g), u_edge=torch.rand(Eb, 6, generator=g), z_pos=torch.randn(Nl, 3, generator=g))
        with torch.no_grad():
            pn, pp, pe, _ = ref(state["h_node"], state["pos"], b["batch_node"], state["h_edge"], b["edge_index"],
                                b["batch_edge"], t, ph["x"], ph["pos"], ph["norm"], ph["batch"])
            ln = ref.node_transition.q_v_posterior(torch.log_softmax(pn, -1), state["log_node"], t, b["batch_node"], v0_prob=True)
            le = ref.edge_transition.q_v_posterior(torch.log_softmax(pe, -1), state["log_edge"], t, b["batch_edge"], v0_prob=True)
            o_rand, o_randn = torch.rand_like, torch.randn_like
            try:
                torch.rand_like = lambda x: d["u_node"]
                nc = rc.log_sample_categorical(ln)
                torch.rand_like = lambda x: d["u_edge"]
                ec = rc.log_sample_categorical(le)
                torch.randn_like = lambda x: d["z_pos"]
                xp = ref.pos_transition.get_prev_from_recon(x_t=state["pos"], x_recon=pp, t=t, batch=b["batch_node"])
            finally:
                torch.rand_like, torch.randn_like = o_rand, o_randn
        state = dict(h_node=ref.node_transition.onehot_encode(nc), pos=xp, h_edge=ref.edge_transition.onehot_encode(ec),
                     log_node=ln, log_edge=le)
        draws.append(d)
        outs.append(dict(pred_node=pn, pred_pos=pp, pred_edge=pe, node_cls=nc, edge_cls=ec, pos=xp, log_node=ln, log_edge=le))
    return dict(seed=seed, knn_margin=margin, steps=list(steps), init=init, draws=draws, outs=outs)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref, sd, rc = reference_model(0)
    meta = dict(state_dict_digest=state_dict_digest(sd), weight_seed=0, torch=torch.__version__,
                keys=sorted((k, tuple(v.shape)) for k, v in sd.items()))
    torch.save(meta, os.path.join(GOLD, "meta.pt"))
    torch.save(forward_fixture(ref, 3, 4, (9, 14), [999, 500, 17, 0], sd=sd), os.path.join(GOLD, "forward_small.pt"))
    torch.save(forward_fixture(ref, 5, 2, 30, [700, 3], stages=False, sd=sd), os.path.join(GOLD, "forward_n30.pt"))
    torch.save(forward_fixture(ref, 7, 2, (20, 26), [250, 900], n_ex=45, stages=False, sd=sd), os.path.join(GOLD, "forward_ex.pt"))
    torch.save(transition_fixture(ref, rc, 21), os.path.join(GOLD, "transition.pt"))
    torch.save(graph_fixture(), os.path.join(GOLD, "graph.pt"))
    torch.save(reverse_steps_fixture(ref, rc, 9, sd=sd), os.path.join(GOLD, "reverse_steps.pt"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    a = sys.argv[1:]
    big_main() if "big" in a else train_main() if "train" in a else phores_main() if "phores" in a else main()
