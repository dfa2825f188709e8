"""Eager (non-graph) reverse steps of a synthetic workload inside a cudaProfilerStart/Stop range; meant to be wrapped by
    ncu --profile-from-start off ...  python tools/profile_step.py [molecules] [steps] [atoms] [phore features] [guidance 0|1]
(one untimed warm-up step runs before the range: module loading, cudaFuncSetAttribute)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phoregen_b200.diffusion import PhoreDiff, TrajectorySampler
from phoregen_b200.synthetic import synthetic_batch
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
G, steps, atoms, p, guide = arg(1, 1024), arg(2, 2), arg(3, 30), arg(4, 0), arg(5, 0)
dev = torch.device("cuda:0")
model = PhoreDiff(MODEL_CONFIG, "zinc_300")
model.load_state_dict(random_state_dict(model, 0), strict=True)
model = model.to(dev).eval()
b = synthetic_batch(2032, G, n_atoms=atoms, p_choices=(p,) if p else (6, 7, 8))
opts = [dict(type="atom_prox", min_d=1.2, max_d=1.9), dict(type="center_prox")] if guide else None
smp = TrajectorySampler(model, None, G, dev, ligand_num_atoms=b["num_atoms"], save_traj=False, seed=2032, use_cuda_graph=False,
                        phore_batch=b["phore"], guidance=opts)
smp.run(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
smp.run(steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", smp.plan.launches)
