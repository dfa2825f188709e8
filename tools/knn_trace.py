"""Dump the phase stamps of the tcgen05 kNN attention kernel, key pass (needs a build with PG_NVCC_EXTRA=-DPG_TRIP_TRACE)."""
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
b = synthetic_batch(2032, 1024, n_atoms=30)
smp = TrajectorySampler(model, None, 1024, dev, ligand_num_atoms=b["num_atoms"], save_traj=False, seed=1, use_cuda_graph=False, phore_batch=b["phore"])
smp.run(1); torch.cuda.synchronize()
buf = np.zeros(2 * 64 * 16, dtype=np.int64)
fn = _lib.lib.pg_debug_knn_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
t = buf.reshape(2, 64, 16)
base = t[1, 8, 0]
for tile in range(8, 16):
    for role, name in ((1, "A"), (0, "R")):
        row = t[role, tile]
        print(f"tile {tile} {name}: " + " ".join(f"{(int(v) - int(base)) if v else -1:>7d}" for v in row[:15]))
    print()

# per row warp (warp = 4 * channel quarter + lane quarter): loop top | PRE wait done | HID arrival | post-processing done
wb = np.zeros(16 * 64 * 4, dtype=np.int64)
fn2 = getattr(_lib.lib, "pg_debug_knn_wtrace", None)
if fn2 is not None:
    fn2.restype = ctypes.c_int; fn2.argtypes = [ctypes.c_void_p]
    assert fn2(wb.ctypes.data) == 0
    w = wb.reshape(16, 64, 4)
    for tile in range(10, 14):
        print(f"tile {tile}: per-warp stamps relative to warp 0's loop top")
        b0 = int(w[0, tile, 0])
        for wi in range(16):
            print(f"  warp {wi:2d} (wq {wi & 3}, cq {wi >> 2}): " + " ".join(f"{int(v) - b0:>7d}" for v in w[wi, tile]))
