"""Measurement aid: is the edge GEMM (write-dominated, 512 B read + 128*nt*4 B written per row) at the HBM write limit?
Times pg_gemm_k128 (tcgen05 kernel) at the config[1] edge count for several output widths next to plain device
write / copy baselines of the same byte counts (torch fill_ / copy_), all with CUDA events after warm-up."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from phoregen_b200._lib import lib, check
from phoregen_b200.weights import bf16_tiles64

dev = torch.device("cuda:0")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
M = int(os.environ.get("M", 1024 * 30 * 29))


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
A = torch.randn(M, 128, device=dev)
for nt in [int(x) for x in os.environ.get("NT", "1,2,3,5,7").split(",")]:
    N = 128 * nt
    Wt = rng.normal(size=(128, N)).astype(np.float32) / 11
    Wbf = torch.from_numpy(bf16_tiles64(Wt)).to(dev)
    bias = torch.zeros(N, device=dev)
    C = torch.empty(M, N, device=dev)
    ms = timed(lambda: check(lib.pg_gemm_k128(0, 0, M, P(A), 128, None, 128, None, None, None, None, P(Wbf), P(bias), None, 0, P(C), N, nt, st), "gemm"))
    wr = M * N * 4
    ms_fill = timed(lambda: C.fill_(1.0))
    src = torch.empty(M, N // 2 if N > 128 else N, device=dev)
    print(f"nt={nt} N={N}: gemm {ms:.3f} ms  ({(wr + M * 512) / ms / 1e6:.0f} GB/s incl. A read, {wr / ms / 1e6:.0f} GB/s written)   "
          f"fill_ of C {ms_fill:.3f} ms ({wr / ms_fill / 1e6:.0f} GB/s)", flush=True)
    del C, src
if os.environ.get("NT"): sys.exit(0)
big = torch.empty(1 << 30, device=dev, dtype=torch.float32)       # 4 GiB
half = big[: 1 << 29]
ms = timed(lambda: big.fill_(0.0)); print(f"fill_ 4 GiB: {ms:.3f} ms  {big.numel() * 4 / ms / 1e6:.0f} GB/s")
ms = timed(lambda: big[1 << 29:].copy_(half)); print(f"copy 2 GiB -> 2 GiB: {ms:.3f} ms  {big.numel() * 4 / ms / 1e6:.0f} GB/s (read + write)")
ms = timed(lambda: half.sum()); print(f"sum over 2 GiB: {ms:.3f} ms  {half.numel() * 4 / ms / 1e6:.0f} GB/s read")
