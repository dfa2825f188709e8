"""Time the per-molecule result views at config[1] size on the CPU: reference unbatch_data + decode_data (imported
unmodified through the shims when /root/reference is mounted) against phoregen_b200.results."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_cpu_results import _fake_results, _reference_functions
from phoregen_b200 import results as R

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
res = _fake_results(0, [30] * G, T=1)
ref_unbatch, ref_decode = _reference_functions()
t0 = time.perf_counter()
mols = ref_unbatch(res, G)
dec = [ref_decode(m["pred"], m["edge_index"]) for m in mols]
t1 = time.perf_counter()
mine = R.decode_batch(res, G)
t2 = time.perf_counter()
mols2 = R.unbatch_data(res, G)
t3 = time.perf_counter()
print(f"G={G}: reference unbatch_data + decode_data {t1 - t0:.3f} s; results.decode_batch {t2 - t1:.3f} s; results.unbatch_data {t3 - t2:.3f} s")
