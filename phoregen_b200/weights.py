"""Pack a reference-layout state_dict (641 keys, SURVEY.md §8(b) "Checkpoint") into the fp32 blob the CUDA
library reads.  Pure re-layout plus the exact algebraic folds described in DESIGN.md ("Factorisation"):
the first Linear of each MLP is split along its concatenated input (reference uni_denoiser.py:43-46,141-147,
190-193), the one-hot edge type (x) smearing product (common.py:156-163, uni_denoiser.py:270-271) becomes four
20x128 weight slices, and `dire_embedding` (uni_denoiser.py:257-258,279) is folded into the first Linear.
Folds are done in float64 and rounded once to fp32.
"""
import numpy as np
import torch

from . import _lib

SUBS = {"nk": "node_layer_with_edge", "nb": "node_layer_with_bond", "tr": "bond_layer",
        "pk": "pos_layer_with_edge", "pb": "pos_layer_with_bond"}
_FN = {"nk": ("hk_func", "hv_func", "hq_func"), "nb": ("hk_func", "hv_func", "hq_func"),
       "tr": ("hk_func", "hv_func", "hq_func"), "pk": ("xk_func", "xv_func", "xq_func"),
       "pb": ("xk_func", "xv_func", "xq_func")}


def _mlp(sd, prefix):
    g = lambda k: sd[prefix + k].detach().double().cpu().numpy()
    return dict(W1=g(".net.0.weight"), b1=g(".net.0.bias"), g=g(".net.1.weight"), b=g(".net.1.bias"),
                W2=g(".net.3.weight"), b2=g(".net.3.bias"))


def _knn_tables(W1, Wde):
    """[4][24][128]: rows 0-19 smear slice of the edge type, 20 the type column, 21-23 dire (folded)."""
    tab = np.zeros((4, 24, 128))
    dire = W1[:, 84:93] @ Wde                       # [128,3]
    for t in range(4):
        tab[t, :20] = W1[:, t * 20:(t + 1) * 20].T
        tab[t, 20] = W1[:, 80 + t]
        tab[t, 21:24] = dire.T
    return tab


def pack_state_dict(sd):
    """-> dict slot name -> float64 ndarray (flattened later)."""
    out = {}
    f = lambda k: sd[k].detach().double().cpu().numpy()
    out["G.node_emb_t"] = f("node_embedder.weight").T
    out["G.edge_emb_t"] = f("edge_embedder.weight").T
    out["G.time_coeff"] = f("time_emb.0.coeff")
    out["G.time_offset"] = f("time_emb.0.offset")
    out["G.ph_emb_wt"] = f("phore_embedding.weight").T
    out["G.ph_emb_b"] = f("phore_embedding.bias")
    # pharmacophore encoder: kv input = [dist(1) | h_dst(128) | h_src(128)]  (models/__init__.py:29-35)
    k, v, q = (_mlp(sd, "phore_encoder." + n) for n in ("hk_func", "hv_func", "hq_func"))
    z = np.zeros(128)
    out["PE.wcat_t"] = np.concatenate([k["W1"][:, 1:129].T, k["W1"][:, 129:257].T, v["W1"][:, 1:129].T,
                                       v["W1"][:, 129:257].T, q["W1"].T], 1)
    out["PE.bcat"] = np.concatenate([k["b1"], z, v["b1"], z, q["b1"]])
    out["PE.wd_k"], out["PE.wd_v"] = k["W1"][:, 0], v["W1"][:, 0]
    for tag, m in (("k", k), ("v", v), ("q", q)):
        out[f"PE.ln{tag}_g"], out[f"PE.ln{tag}_b"] = m["g"], m["b"]
    out["PE.w2q_t"], out["PE.b2q"] = q["W2"].T, q["b2"]
    out["PE.w2k"], out["PE.b2k"], out["PE.w2v"], out["PE.b2v"] = k["W2"], k["b2"], v["W2"], v["b2"]
    # global edge weight MLP (uni_denoiser.py:326,410-415)
    e = _mlp(sd, "denoiser.edge_pred_layer")
    out["G.ew.w1t"], out["G.ew.b1"], out["G.ew.ln_g"], out["G.ew.ln_b"] = e["W1"].T, e["b1"], e["g"], e["b"]
    out["G.ew.w2"] = e["W2"][0]
    out["G.ew.b2"] = np.concatenate([e["b2"], np.zeros(3)])
    out["G.vinf.w1t"], out["G.vinf.b1"] = f("v_inference.0.weight").T, f("v_inference.0.bias")
    out["G.vinf.w2"], out["G.vinf.b2"] = f("v_inference.2.weight"), f("v_inference.2.bias")
    for c, name in enumerate(("atom_mlp", "atom_mlp_1")):          # atom-count heads (diffusion.py:77-88)
        out[f"G.cnt{c}.w1t"], out[f"G.cnt{c}.b1"] = f(name + ".0.weight").T, f(name + ".0.bias")
        out[f"G.cnt{c}.w2"] = f(name + ".2.weight")[0]
        out[f"G.cnt{c}.b2"] = np.concatenate([f(name + ".2.bias"), np.zeros(3)])
    out["G.binf.w1t"], out["G.binf.b1"] = f("bond_inference.0.weight").T, f("bond_inference.0.bias")
    out["G.binf.w2"] = f("bond_inference.2.weight")
    out["G.binf.b2"] = np.concatenate([f("bond_inference.2.bias"), np.zeros(2)])

    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("denoiser.base_block."))
    for l in range(n_layers):
        P = f"denoiser.base_block.{l}."
        L = f"L{l}."
        m = {s: tuple(_mlp(sd, P + SUBS[s] + "." + fn) for fn in _FN[s]) for s in SUBS}
        Wde, bde = f(P + "dire_embedding.weight"), f(P + "dire_embedding.bias")      # [9,3], [9]

        def knn_cols(s):       # kv input = [edge_feat(93) | h_dst(128) | h_src(128)]
            k, v, q = m[s]
            cols, bias = [], []
            for mm in (k, v):
                cols += [mm["W1"][:, 93:221].T, mm["W1"][:, 221:349].T]
                bias += [mm["b1"] + mm["W1"][:, 84:93] @ bde, z]
            return cols + [q["W1"].T], bias + [q["b1"]]

        def bond_cols(s):      # kv input = [h_bond(128) | h_dst(128) | h_src(128)]
            k, v, q = m[s]
            cols, bias = [], []
            for mm in (k, v):
                cols += [mm["W1"][:, 128:256].T, mm["W1"][:, 256:384].T]
                bias += [mm["b1"], z]
            return cols + [q["W1"].T], bias + [q["b1"]]

        tk, tv, tq = m["tr"]   # kv = [h_bond_kj(128) | r_kj(20) | r_ji(20) | a(13) | h_k(128) | h_j(128)], q = [h_bond_ji | h_i]
        c1, b1 = knn_cols("nk")
        c2, b2 = bond_cols("nb")
        c3 = [tk["W1"][:, 181:309].T, tk["W1"][:, 309:437].T, tv["W1"][:, 181:309].T, tv["W1"][:, 309:437].T,
              tq["W1"][:, 128:256].T]
        b3 = [z, tk["b1"], z, tv["b1"], tq["b1"]]
        out[L + "n1.wt"] = np.concatenate(c1 + c2 + c3, 1)
        out[L + "n1.b"] = np.concatenate(b1 + b2 + b3)
        out[L + "e1.wt"] = np.concatenate([m["nb"][0]["W1"][:, :128].T, m["nb"][1]["W1"][:, :128].T,
                                           tk["W1"][:, :128].T, tv["W1"][:, :128].T, tq["W1"][:, :128].T], 1)
        out[L + "e1.b"] = np.zeros(640)
        c4, b4 = knn_cols("pk")
        c5, b5 = bond_cols("pb")
        out[L + "n2.wt"] = np.concatenate(c4 + c5, 1)
        out[L + "n2.b"] = np.concatenate(b4 + b5)
        out[L + "e2.wt"] = np.concatenate([m["pb"][0]["W1"][:, :128].T, m["pb"][1]["W1"][:, :128].T], 1)
        out[L + "e2.b"] = np.zeros(256)
        out[L + "lin.wt"], out[L + "lin.b"] = f(P + "lin_node.weight").T, f(P + "lin_node.bias")
        for s in SUBS:
            k, v, q = m[s]
            S = L + s + "."
            out[S + "lnq_g"], out[S + "lnq_b"], out[S + "w2q_t"], out[S + "b2q"] = q["g"], q["b"], q["W2"].T, q["b2"]
            out[S + "lnk_g"], out[S + "lnk_b"], out[S + "lnv_g"], out[S + "lnv_b"] = k["g"], k["b"], v["g"], v["b"]
            out[S + "w2k"], out[S + "b2k"], out[S + "w2v"], out[S + "b2v"] = k["W2"], k["b2"], v["W2"], v["b2"]
            if s in ("nk", "pk"):
                out[S + "tab_k"], out[S + "tab_v"] = _knn_tables(k["W1"], Wde), _knn_tables(v["W1"], Wde)
            if s == "tr":
                out[S + "wrkj"] = np.concatenate([k["W1"][:, 128:148].T, v["W1"][:, 128:148].T], 1)
                out[S + "wrji"] = np.concatenate([k["W1"][:, 148:168].T, v["W1"][:, 148:168].T], 1)
                out[S + "wa"] = np.concatenate([k["W1"][:, 168:181].T, v["W1"][:, 168:181].T], 1)
        _center_first_linears(out, L)
    # h_bond does not change between the e2 GEMM of a layer and the e1 GEMM of the next: one 896-column weight for both
    for l in range(n_layers - 1):
        out[f"L{l}.e2e1.wt"] = np.concatenate([out[f"L{l}.e2.wt"], out[f"L{l + 1}.e1.wt"]], 1)
    return out


def _center_first_linears(out, L):
    """Every first Linear of the attention MLPs is followed by a LayerNorm over its 128 hidden channels (common.py:99-119
    with norm=True), which subtracts the channel mean of the pre-activation.  That subtraction is linear, so it is applied
    to the weights instead: each 128-column block of the split first Linears (node / edge GEMM blocks and their biases,
    the kNN edge-type tables, the triplet smearing and angle slices) has its mean over the 128 output channels removed
    (fp64).  Every partial product the kernels add up is then mean-free, their sum is the pre-activation minus its mean, and
    the tensor-core attention kernels only need the second moment (sum x^2) for the LayerNorm.  Kernels that still
    subtract the mean (fp32 twins, the GEMM prologue of the query MLPs) see a mean of ~0: same result."""
    def blocks(a):
        a = np.array(a, dtype=np.float64)
        shp = a.shape
        v = a.reshape(shp[:-1] + (shp[-1] // 128, 128))
        return (v - v.mean(axis=-1, keepdims=True)).reshape(shp)
    for name in ("n1.wt", "n1.b", "e1.wt", "e1.b", "n2.wt", "n2.b", "e2.wt", "e2.b", "nk.tab_k", "nk.tab_v", "pk.tab_k", "pk.tab_v",
                 "tr.wrkj", "tr.wrji", "tr.wa"):
        out[L + name] = blocks(out[L + name])


def _bf16_bits(x32):
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x32, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def bf16_split(wt):
    """k-major fp32 weight [128][N] -> fp32-typed carrier of [hi|lo][N][128] bf16 (K-major), hi + lo ~ w to 2^-17.
    This is the B operand of the tcgen05 'bf16x3' contraction (csrc/pg_gemm_tc.cu)."""
    w = np.ascontiguousarray(np.asarray(wt, dtype=np.float64).T, dtype=np.float32)      # [N][128]
    hi = _bf16_bits(w)
    hi_f = (hi.astype(np.uint32) << 16).view(np.float32)
    lo = _bf16_bits(w - hi_f)
    return np.concatenate([hi.reshape(-1), lo.reshape(-1)]).view(np.float32)


def f16_image(wt):
    """k-major fp32 weight [128][N] -> fp32-typed carrier of [N][128] IEEE half values (K-major): the B operand of the
    single-pass fp16 contraction the tensor-core attention kernels use for the key MLP's second Linear."""
    w = np.ascontiguousarray(np.asarray(wt, dtype=np.float64).T, dtype=np.float32)
    return w.astype(np.float16).view(np.uint16).reshape(-1).view(np.float32)


def bf16_tiles64(wt):
    """k-major fp32 weight [128][N] -> the exact shared-memory images the tcgen05 GEMM bulk-copies: for every block of 64
    output columns a 32 KB image [hi|lo][K block 0|1][64 rows x 128 B], rows K-major with the 128-byte swizzle applied
    (16-byte chunk j of row r stored at chunk j ^ (r % 8)); see csrc/pg_gemm_tc.cu and tc::sw128_chunk."""
    w = np.ascontiguousarray(np.asarray(wt, dtype=np.float64).T, dtype=np.float32)      # [N][128]
    N = w.shape[0]
    assert N % 64 == 0 and w.shape[1] == 128
    hi = _bf16_bits(w)
    lo = _bf16_bits(w - (hi.astype(np.uint32) << 16).view(np.float32))
    r = np.arange(64)[:, None]
    j = np.arange(8)[None, :]
    out = np.empty((N // 64, 2, 2, 64, 8, 8), dtype=np.uint16)
    for part, bits in enumerate((hi, lo)):
        t = bits.reshape(N // 64, 64, 2, 8, 8)                 # [tile][row][kb][chunk][8]
        for kb in range(2):
            out[:, part, kb][:, r, j ^ (r % 8)] = t[:, :, kb][:, r, j]
    return out.reshape(-1).view(np.float32)


def angle_slab_image(wa13):
    """Angle slice of the triplet MLPs' first Linear, [13][256] (k | v) -> the shared-memory image of the tcgen05 triplet
    kernel's angle slab (csrc/pg_trip_tc.cu): [mlp][hi|lo][channel half][16 rows][64 bf16], MN-major with the 128-byte
    swizzle (16-byte chunk c of row r stored at chunk c ^ (r % 8)).  Rows 0..10 are the 11 DISTINCT angular features
    [theta, sin(theta * {1,2,3}), sin(theta / {2,3}), cos(theta * {1,2,3}), cos(theta / {2,3})]: the reference encoding
    (common.py:67-87) lists sin(theta) and cos(theta) twice (frequencies 1 and 1/1), so their weight rows are added here
    (fp64).  Rows 11..14 receive the tile's R[j->i] rows at run time, row 15 stays zero."""
    w = np.asarray(wa13, dtype=np.float64)
    assert w.shape == (13, 256)
    rows = np.zeros((16, 256))
    rows[0] = w[0]; rows[1] = w[1] + w[4]; rows[2] = w[2]; rows[3] = w[3]; rows[4] = w[5]; rows[5] = w[6]
    rows[6] = w[7] + w[10]; rows[7] = w[8]; rows[8] = w[9]; rows[9] = w[11]; rows[10] = w[12]
    r32 = rows.astype(np.float32)
    hi = _bf16_bits(r32)
    lo = _bf16_bits(r32 - (hi.astype(np.uint32) << 16).view(np.float32))
    out = np.zeros((2, 2, 2, 16, 8, 8), dtype=np.uint16)                 # [mlp][hl][half][row][chunk][8]
    r = np.arange(16)[:, None]
    c = np.arange(8)[None, :]
    for hl, bits in enumerate((hi, lo)):
        t = bits.reshape(16, 2, 2, 8, 8)                                  # [row][mlp][half][chunk][8]
        for mlp in range(2):
            for half in range(2):
                out[mlp, hl, half][r, c ^ (r % 8)] = t[:, mlp, half][r, c]
    return out.reshape(-1).view(np.float32)


def build_blob(sd):
    """-> (fp32 numpy blob, int64 offsets in floats) following the library's own slot table."""
    packed = pack_state_dict(sd)
    for name in [k for k in packed if k.endswith((".wt", "w2q_t", "wcat_t", "w1t")) and k != "G.ew.w1t" and not k.startswith("G.cnt")]:
        packed[name + ".bf"] = bf16_tiles64(packed[name])
    # Tensor-core attention kernels: second Linear as bf16 hi/lo images.  Where every LayerNorm gain of the MLP is positive,
    # relu(g*xn + b) = g * relu(xn + b/g), so g is folded into the columns of W2 (fp64) and the kernel only needs b/g
    # ("<S>.ln{k,v}_bf" holds b/g, or plain b when not folded; "<S>.fold" = [fold_k, fold_v, 0, 0]).  The fp32 twins keep
    # using the unfolded "<S>.w2k" / "<S>.lnk_b".
    for S in sorted({k[:-len("lnk_g")] for k in packed if k.endswith((".tr.lnk_g", ".nb.lnk_g", ".pb.lnk_g", ".nk.lnk_g", ".pk.lnk_g"))}):
        flags = np.zeros(4)
        for i, kv in enumerate("kv"):
            g = np.asarray(packed[S + f"ln{kv}_g"], dtype=np.float64)
            b = np.asarray(packed[S + f"ln{kv}_b"], dtype=np.float64)
            w2 = np.asarray(packed[S + f"w2{kv}"], dtype=np.float64)              # [out][in]
            fold = bool(np.all(g > 1e-3))
            flags[i] = 1.0 if fold else 0.0
            packed[S + f"ln{kv}_bf"] = b / g if fold else b
            packed[S + f"w2{kv}.bf"] = bf16_split((w2 * g[None, :] if fold else w2).T)   # [out][in]; pos value head: 16 outputs
            if kv == "k":
                packed[S + "w2k.h"] = f16_image((w2 * g[None, :] if fold else w2).T)
        packed[S + "fold"] = flags
    for name in [k for k in packed if k.endswith((".nk.tab_k", ".nk.tab_v", ".pk.tab_k", ".pk.tab_v"))]:
        packed[name + ".bf"] = bf16_split(np.asarray(packed[name]).reshape(96, 128))   # [type*24 + feat][128] -> [hi|lo][128][96]
    for name in [k for k in packed if k.endswith(".tr.wa")]:
        packed[name + ".bf"] = angle_slab_image(packed[name])
    table = _lib.slot_table()
    offsets = np.zeros(len(table), dtype=np.int64)
    chunks, cur = [], 0
    for i, (name, numel) in enumerate(table):
        if name not in packed:
            raise KeyError(f"weight packer does not produce slot {name!r}")
        raw = packed[name]
        arr = raw.reshape(-1) if raw.dtype == np.float32 else np.ascontiguousarray(raw, dtype=np.float64).reshape(-1).astype(np.float32)
        if arr.size != numel:
            raise ValueError(f"slot {name}: expected {numel} values, packed {arr.size}")
        pad = (-arr.size) % 4
        offsets[i] = cur
        chunks.append(arr)
        if pad:
            chunks.append(np.zeros(pad, dtype=np.float32))
        cur += arr.size + pad
    extra = set(packed) - {n for n, _ in table}
    if extra:
        raise KeyError(f"packer produced unknown slots: {sorted(extra)[:5]}")
    return np.concatenate(chunks), offsets


def blob_to_device(sd, device):
    blob, offsets = build_blob(sd)
    return torch.from_numpy(blob).to(device), offsets
