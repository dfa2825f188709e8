"""Debug aid: forward of the reverse_steps fixture, error vs golden (use env PG_TRIP=fp32 / PG_GEMM=simt to bisect)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_model, load_golden, report
from oracle import phoregen_oracle as O
from phoregen_b200.engine import BatchPlan
dev = torch.device("cuda:0")
m, sd = build_model(dev)
pm = m.packed(dev)
f = load_golden("reverse_steps.pt")
b = O.synthetic_batch(f["seed"], 3, n_atoms=(8, 11))
print("num_atoms", b["num_atoms"].tolist(), "phore", torch.bincount(b["phore"]["batch"]).tolist())
ph = {k: v.to(dev) for k, v in b["phore"].items()}
for order in (0, 2):
    if order == 0:
        plan = BatchPlan(b["num_atoms"].numpy(), torch.bincount(b["phore"]["batch"]).numpy(), dev, edge_order=0)
    else:
        plan = BatchPlan(b["num_atoms"].numpy(), torch.bincount(b["phore"]["batch"]).numpy(), dev, ref_edge_index=b["edge_index"].to(dev))
    hp = plan.phore_encode(pm, ph["x"], ph["pos"])
    st = {k: v.to(dev).clone() for k, v in f["init"].items()}
    t = torch.full((3,), 999, dtype=torch.long, device=dev)
    for rep in range(2):
        pn, pp, pe = plan.phorediff_forward(pm, st["h_node"], st["pos"], st["h_edge"], t, hp, ph["pos"], ph["norm"])
        w = f["outs"][0]
        print(f"order {order} rep {rep}: node", report(pn, w["pred_node"]), "| edge", report(pe, w["pred_edge"]), "| pos", report(pp, w["pred_pos"]))
