"""Debug aid: tcgen05 GEMM vs fp64 reference for several shapes / prologues."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from phoregen_b200._lib import lib, check
from phoregen_b200.weights import bf16_tiles64
dev = torch.device("cuda:0")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(0)
for M in (22, 53, 300, 40000):
    for nt in (5, 10, 15):
        for pro in (0, 1, 2):
            N = 128 * nt
            lda = 256
            A = torch.from_numpy(rng.normal(size=(M, lda)).astype(np.float32)).to(dev)
            A2 = torch.from_numpy(rng.normal(size=(max(M, 64), 128)).astype(np.float32)).to(dev)
            gidx = torch.from_numpy(rng.integers(0, A2.shape[0], size=M).astype(np.int32)).to(dev)
            g = torch.from_numpy(rng.normal(size=128).astype(np.float32)).to(dev); b = torch.from_numpy(rng.normal(size=128).astype(np.float32)).to(dev)
            Wt = rng.normal(size=(128, N)).astype(np.float32) / 11
            bias = torch.from_numpy(rng.normal(size=N).astype(np.float32)).to(dev)
            Wt_d = torch.from_numpy(Wt).to(dev); Wbf = torch.from_numpy(bf16_tiles64(Wt)).to(dev)
            x = A[:, :128].double()
            if pro == 1: x = x + A2[:M].double()
            if pro == 2:
                x = x + A2[gidx.long()].double()
                x = torch.relu(torch.nn.functional.layer_norm(x, (128,), g.double(), b.double(), 1e-5))
            want = x @ Wt_d.double() + bias.double()
            outs = []
            for impl in (0, 1):
                C = torch.full((M, N), float("nan"), device=dev)
                check(lib.pg_gemm_k128(impl, pro, M, P(A), lda, P(A2) if pro else None, 128, P(gidx) if pro == 2 else None, P(g), P(b), P(Wt_d), P(Wbf),
                                       P(bias), None, 0, P(C), N, nt, st), "gemm")
                torch.cuda.synchronize()
                outs.append((C.double() - want).abs().max().item())
            flag = "  <-- BAD" if not (outs[0] < 1e-3) else ""
            print(f"M={M:6d} N={N:4d} pro={pro}: tc err {outs[0]:.2e}  simt err {outs[1]:.2e}{flag}")
