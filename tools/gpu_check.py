"""Stage-by-stage GPU-vs-golden report (debug aid; the graded checks live in tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_model, load_golden, report
from oracle import phoregen_oracle as O
from phoregen_b200.engine import BatchPlan
from phoregen_b200.testing import state_dict_digest

dev = torch.device("cuda:0")
model, sd = build_model(dev)
meta = load_golden("meta.pt")
print("digest ok:", meta["state_dict_digest"] == state_dict_digest(sd))
pm = model.packed(dev)

g = load_golden("graph.pt")
plan = BatchPlan(g["num_atoms"].numpy(), g["num_phore"].numpy(), dev, edge_order=0)
ei, eb = plan.bond_edges()
print("bond edges:", torch.equal(ei.cpu(), g["edge_index"]), torch.equal(eb.cpu(), g["edge_batch"]))
k32 = plan.knn_graph(g["x"].to(dev), 0)
print("knn32:", torch.equal(k32.cpu(), g["knn32"]), k32.shape, g["knn32"].shape)
k3 = plan.knn_graph(g["x"].to(dev), 1)
print("knn3:", torch.equal(k3.cpu(), g["knn3"]), k3.shape, g["knn3"].shape)
tr = plan.triplets()
print("triplets:", [torch.equal(a.cpu(), b) for a, b in zip(tr, g["triplets"])])

for name in ("forward_small.pt", "forward_n30.pt", "forward_ex.pt"):
    f = load_golden(name)
    b = O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=f["n_atoms"], n_ex=f["n_ex"])
    ph = b["phore"]
    to = lambda t: t.to(dev)
    t = torch.tensor(f["times"], dtype=torch.long, device=dev)
    torch.cuda.synchronize(); t0 = time.time()
    out = model(to(b["h_node"]), to(b["pos"]), to(b["batch_node"]), to(b["h_edge"]), to(b["edge_index"]), to(b["batch_edge"]), t,
                to(ph["x"]), to(ph["pos"]), to(ph["norm"]), to(ph["batch"]))
    torch.cuda.synchronize()
    print(name, f"{time.time()-t0:.3f}s")
    print("  pred_node", report(out[0], f["pred_node"]))
    print("  pred_pos ", report(out[1], f["pred_pos"]))
    print("  pred_edge", report(out[2], f["pred_edge"]))
    print("  count    ", report(out[3][0], f["count_l"]), report(out[3][1], f["count_u"]))
    if "h_phore_emb" in f:
        na = torch.bincount(b["batch_node"]).numpy(); npn = torch.bincount(ph["batch"]).numpy()
        plan = BatchPlan(na, npn, dev, ref_edge_index=to(b["edge_index"]))
        hp = plan.phore_encode(pm, to(ph["x"]), to(ph["pos"]))
        print("  h_phore_emb", report(hp, f["h_phore_emb"]))

big = load_golden("forward_big.pt")["cases"]
for name, f in big.items():
    b = O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=f["n_atoms"], p_choices=f["p_choices"], pos_scale=f["pos_scale"])
    ph = b["phore"]
    to = lambda t: t.to(dev)
    out = model(to(b["h_node"]), to(b["pos"]), to(b["batch_node"]), to(b["h_edge"]), to(b["edge_index"]), to(b["batch_edge"]),
                torch.tensor(f["times"], dtype=torch.long, device=dev), to(ph["x"]), to(ph["pos"]), to(ph["norm"]), to(ph["batch"]))
    print("big", name, "| node", report(out[0], f["pred_node"]), "| pos", report(out[1], f["pred_pos"]), "| edge", report(out[2], f["pred_edge"]))

# denoiser drop-in outputs (h, h_bond, x after the 6 layers) against the reference's hooked tensors
f = load_golden("forward_small.pt")
b = O.synthetic_batch(f["seed"], f["n_graphs"], n_atoms=f["n_atoms"])
ph = b["phore"]
stages = []
O.phorediff_forward(sd, b["h_node"], b["pos"], b["batch_node"], b["h_edge"], b["edge_index"], b["batch_edge"], torch.tensor(f["times"]),
                    ph["x"], ph["pos"], ph["norm"], ph["batch"], stages=stages)
s0 = stages[0]
to = lambda t: t.to(dev)
out = model.denoiser(h=to(s0["h_all"]), x=to(s0["pos_all"]), group_idx=None, bond_index=to(s0["bond_index_in_all"]), h_bond=to(s0["h_edge"]),
                     mask_ligand=to(s0["mask_ligand"]), mask_ligand_atom=to(s0["mask_ligand"]), batch=to(s0["batch_all"]),
                     phore_norm=to(ph["norm"]), packed=pm)
h5, hb5, x5 = f["layer5"]
print("layer5 | h", report(out["h"], h5), "| h_bond", report(out["h_bond"], hb5), "| x", report(out["x"], x5))
