"""Run a few eager (non-graph) reverse steps of the config[1] workload; meant to be wrapped by ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phoregen_b200.diffusion import PhoreDiff, TrajectorySampler
from phoregen_b200.synthetic import synthetic_batch
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
model = PhoreDiff(MODEL_CONFIG, "zinc_300")
model.load_state_dict(random_state_dict(model, 0), strict=True)
model = model.to(dev).eval()
b = synthetic_batch(2032, G, n_atoms=30)
smp = TrajectorySampler(model, None, G, dev, ligand_num_atoms=b["num_atoms"], save_traj=False, seed=2032, use_cuda_graph=False,
                        phore_batch=b["phore"])
smp.run(steps)
torch.cuda.synchronize()
print("launches", smp.plan.launches)
