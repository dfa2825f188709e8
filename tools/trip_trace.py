"""Dump the phase stamps of the tcgen05 triplet kernel (needs a build with PG_NVCC_EXTRA=-DPG_TRIP_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from phoregen_b200 import _lib
from phoregen_b200.diffusion import PhoreDiff, TrajectorySampler
from phoregen_b200.synthetic import synthetic_batch
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict
dev = torch.device("cuda:0")
model = PhoreDiff(MODEL_CONFIG, "zinc_300"); model.load_state_dict(random_state_dict(model, 0)); model = model.to(dev).eval()
b = synthetic_batch(2032, 256, n_atoms=30)
smp = TrajectorySampler(model, None, 256, dev, ligand_num_atoms=b["num_atoms"], save_traj=False, seed=1, use_cuda_graph=False, phore_batch=b["phore"])
smp.run(1); torch.cuda.synchronize()
buf = np.zeros(3 * 64 * 16, dtype=np.int64)
fn = _lib.lib.pg_debug_trip_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
t = buf.reshape(3, 64, 16)
base = t[2, 8, 0]
names = {0: "K", 1: "V", 2: "M"}
for tile in range(8, 20):
    for role in (2, 0, 1):
        row = t[role, tile]
        print(f"tile {tile} {names[role]}: " + " ".join(f"{(int(v) - int(base)) if v else -1:>7d}" for v in row[:14]))
    print()
