"""Time `PhoreDiff.compute_loss` (forward-only training objective) at the config[1] shape: molecules per second of one
validation pass (noise draws + CUDA forward + loss terms), CUDA events around 5 calls after 2 warm-up calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phoregen_b200.diffusion import PhoreDiff
from phoregen_b200.synthetic import synthetic_batch
from phoregen_b200.testing import MODEL_CONFIG, random_state_dict, training_batch_from_synthetic

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
model = PhoreDiff(MODEL_CONFIG, "zinc_300"); model.load_state_dict(random_state_dict(model, 0), strict=True); model = model.to(dev).eval()
data = training_batch_from_synthetic(synthetic_batch(2032, G, n_atoms=30, edge_order="training")).to(dev)
for _ in range(2):
    model.compute_loss(data)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    total, d = model.compute_loss(data)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"compute_loss: {ms:.1f} ms per batch of {G} molecules = {G / ms * 1e3:.0f} molecules/s (loss {d['loss']:.3f})")
