"""Forward outputs of the tensor-core path against the fp32 FFMA twins at given molecule sizes (separate processes, because
the kernel-selection switches are read once per process):
    python tools/ab_forward.py run out.pt 33 34 100 128      # current environment
    python tools/ab_forward.py cmp a.pt b.pt
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

if sys.argv[1] == "run":
    from phoregen_b200.diffusion import PhoreDiff
    from phoregen_b200.synthetic import synthetic_batch
    from phoregen_b200.testing import MODEL_CONFIG, random_state_dict
    dev = torch.device("cuda:0")
    model = PhoreDiff(MODEL_CONFIG, "zinc_300")
    model.load_state_dict(random_state_dict(model, 0), strict=True)
    model = model.to(dev).eval()
    out = {}
    for n in [int(v) for v in sys.argv[3:]]:
        b = synthetic_batch(77 + n, 3, n_atoms=n)
        ph = b["phore"]
        t = torch.tensor([900, 400, 20])
        to = lambda x: x.to(dev)
        o = model(to(b["h_node"]), to(b["pos"]), to(b["batch_node"]), to(b["h_edge"]), to(b["edge_index"]), to(b["batch_edge"]), to(t),
                  to(ph["x"]), to(ph["pos"]), to(ph["norm"]), to(ph["batch"]))
        out[n] = [x.cpu() for x in o[:3]]
    torch.save(out, sys.argv[2])
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for n in sorted(set(a) & set(b)):
        worst = 0.0
        for x, y in zip(a[n], b[n]):
            worst = max(worst, float(((x - y).abs() / (1e-4 + 1e-3 * y.abs())).max()))
        print(f"n={n}: worst |tc - fp32| / tolerance = {worst:.3f}", "OK" if worst < 1.0 else "FAIL")
