"""Run the same eager trajectory several times and count bitwise mismatches (debug aid for races in the pipelined kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phoregen_b200.diffusion import PhoreDiff, TrajectorySampler
from phoregen_b200.synthetic import synthetic_batch
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict
dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 320
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
model = PhoreDiff(MODEL_CONFIG, "zinc_300"); model.load_state_dict(random_state_dict(model, 0), strict=True); model = model.to(dev).eval()
lo, hi = (int(x) for x in os.environ.get("DC_ATOMS", "26,33").split(","))      # DC_ATOMS=36,48 exercises the chunked (n > 33) kernels
b = synthetic_batch(2039, G, n_atoms=(lo, hi))
ref = None; bad = 0
for r in range(reps):
    use_graph = bool(int(os.environ.get("DC_GRAPH", "0"))) and (r % 2 == 1)
    s = TrajectorySampler(model, None, G, dev, ligand_num_atoms=b["num_atoms"], save_traj=False, seed=7, use_cuda_graph=use_graph, phore_batch=b["phore"])
    s.run(steps); torch.cuda.synchronize()
    out = (s.pos.clone(), s.node_cls.clone(), s.edge_cls.clone())
    if ref is None: ref = out
    else:
        same = all(torch.equal(a, c) for a, c in zip(ref, out))
        if not same:
            bad += 1
            print("rep", r, "graph" if use_graph else "eager", "differs: max |dpos| =", float((ref[0] - out[0]).abs().max()), "n pos", int((ref[0] != out[0]).sum()),
                  "n node", int((ref[1] != out[1]).sum()), "n edge", int((ref[2] != out[2]).sum()), flush=True)
print("mismatching repetitions:", bad, "of", reps - 1, flush=True)
