"""Host-side handles over the C ABI: packed model, batch plan, and the per-call wrappers.

PyTorch is used here for device memory, streams and (elsewhere) torch.distributed only; every kernel on the
hot path lives in libphoregen_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .weights import blob_to_device


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return ctypes.c_void_p(t.data_ptr())


def _stream(device=None):
    """Current stream OF `device` (not of the process's current device): every launch of a plan goes to the stream of the
    plan's own GPU, whatever device is current in the caller."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on_device(method):
    """Run a BatchPlan method with the plan's device current (kernel launches, cudaFuncSetAttribute and event calls inside
    the library act on the current device) and check that tensor arguments live there."""
    import functools

    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        for t in list(args) + list(kwargs.values()):
            if isinstance(t, torch.Tensor) and t.is_cuda and t.device != self.device:
                raise ValueError(f"{method.__name__}: tensor on {t.device}, plan on {self.device}")
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapped


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _alloc(shape, dtype, device):
    try:
        return torch.empty(shape, dtype=dtype, device=device)
    except torch.OutOfMemoryError as e:  # callers pattern-match this phrase (sample_all.py:96, run/run.py:145)
        raise RuntimeError(f"CUDA out of memory while allocating phoregen_b200 work space: {e}") from e


class PackedModel:
    """Weights of a PhoreDiff state_dict packed for the CUDA kernels."""

    def __init__(self, state_dict, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PhoreGenLibraryError("phoregen_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.blob, offsets = blob_to_device(state_dict, self.device)
        self._offsets = offsets
        h = ctypes.c_void_p()
        check(lib.pg_model_create(ctypes.byref(h), _ptr(self.blob),
                                  offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(offsets)), "pg_model_create")
        self.handle = h
        g = lambda k: state_dict[k].detach().to(self.device, torch.float32).contiguous()
        self.tables = {k: g(k) for k in (
            "pos_transition.coef_x0", "pos_transition.coef_xt", "pos_transition.std",
            "node_transition.q_mats", "node_transition.transpopse_q_onestep_mats",
            "edge_transition.q_mats", "edge_transition.transpopse_q_onestep_mats") if k in state_dict}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.pg_model_destroy(h)
            self.handle = None


class BatchPlan:
    """Static topology + work space of one batch (see pg_plan_create in include/phoregen_b200.h)."""

    def __init__(self, num_atoms, num_phore, device, edge_order=0, ref_edge_index=None):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        na = np.ascontiguousarray(np.asarray(num_atoms, dtype=np.int32))
        npn = np.ascontiguousarray(np.asarray(num_phore, dtype=np.int32))
        assert na.shape == npn.shape and na.ndim == 1
        self.num_atoms, self.num_phore, self.G = na, npn, int(na.size)
        pa = na.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        pp = npn.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        nbytes = int(lib.pg_plan_workspace_bytes(self.G, pa, pp))
        if nbytes < 0:
            check(nbytes, "pg_plan_workspace_bytes")
        self.workspace = _alloc((nbytes + 256,), torch.uint8, self.device)
        base = self.workspace.data_ptr()
        aligned = (base + 255) & ~255
        self.workspace_bytes = nbytes
        if ref_edge_index is not None:
            ref_edge_index = ref_edge_index.to(self.device, torch.int64).contiguous()
            edge_order = 2
        self.edge_order = edge_order
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.pg_plan_create(ctypes.byref(h), self.G, pa, pp, edge_order, _ptr(ref_edge_index),
                                     ctypes.c_void_p(aligned), nbytes, _stream(self.device)), "pg_plan_create")
        self.handle = h
        self.Nl = int(lib.pg_plan_num_ligand_atoms(h))
        self.P = int(lib.pg_plan_num_phore_nodes(h))
        self.N = self.Nl + self.P
        self.Eb = int(lib.pg_plan_num_bond_edges(h))
        self.Ek = int(lib.pg_plan_num_knn_edges(h))
        self.E3 = int(lib.pg_plan_num_triplets(h))
        self._lig_graph_ptr = lib.pg_plan_ligand_graph(h)
        self._edge_graph_ptr = lib.pg_plan_edge_graph(h)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.pg_plan_destroy(h)
            self.handle = None

    @property
    def launches(self):
        return int(lib.pg_plan_kernel_launches(self.handle))

    KERNEL_CLASSES = ("gemm", "knn_attn", "bond_attn", "trip", "knn_graph", "other")

    @_on_device
    def timing(self, on):
        check(lib.pg_plan_timing_enable(self.handle, int(bool(on))), "pg_plan_timing_enable")

    @_on_device
    def read_timing(self):
        """-> {class: (total_ms, launches)} measured with CUDA events on the launching stream."""
        out = {}
        for i, name in enumerate(self.KERNEL_CLASSES):
            ms, n = ctypes.c_double(), ctypes.c_int64()
            check(lib.pg_plan_timing_read(self.handle, i, ctypes.byref(ms), ctypes.byref(n)), "pg_plan_timing_read")
            out[name] = (ms.value, n.value)
        return out

    # ---- graph artefacts (G1, B1, K1) ------------------------------------------------------
    @_on_device
    def bond_edges(self):
        ei = _alloc((2, self.Eb), torch.int64, self.device)
        eb = _alloc((self.Eb,), torch.int64, self.device)
        check(lib.pg_plan_export_bond_edges(self.handle, _ptr(ei), _ptr(eb), _stream(self.device)), "pg_plan_export_bond_edges")
        return ei, eb

    @_on_device
    def triplets(self):
        outs = [_alloc((self.E3,), torch.int64, self.device) for _ in range(5)]
        check(lib.pg_plan_export_triplets(self.handle, *[_ptr(o) for o in outs], _stream(self.device)), "pg_plan_export_triplets")
        return outs

    @_on_device
    def knn_graph(self, x, mode=0):
        """mode 0: k=32 joint graph in context numbering; mode 1: k=3 ligand-only graph in ligand numbering."""
        x = _f32(x)
        assert x.shape == (self.N, 3)
        if mode == 0:
            E = self.Ek
        else:
            E = int(sum(int(n) * min(3, int(n) - 1) for n in self.num_atoms))
        ei = _alloc((2, E), torch.int64, self.device)
        check(lib.pg_knn_graph(self.handle, _ptr(x), mode, _ptr(ei), _stream(self.device)), "pg_knn_graph")
        return ei

    # ---- forward passes ---------------------------------------------------------------------
    @_on_device
    def phore_encode(self, model, h_phore, pos_phore):
        h_phore, pos_phore = _f32(h_phore), _f32(pos_phore)
        assert h_phore.shape == (self.P, 18) and pos_phore.shape == (self.P, 3)
        out = _alloc((self.P, 128), torch.float32, self.device)
        check(lib.pg_phore_encode(model.handle, self.handle, _ptr(h_phore), _ptr(pos_phore), _ptr(out), _stream(self.device)),
              "pg_phore_encode")
        return out

    @_on_device
    def denoiser_forward(self, model, h, x, h_bond, phore_norm):
        h, x, h_bond, phore_norm = _f32(h), _f32(x), _f32(h_bond), _f32(phore_norm)
        assert h.shape == (self.N, 128) and x.shape == (self.N, 3) and h_bond.shape == (self.Eb, 128)
        assert phore_norm.shape == (self.P, 3)
        ho, xo, bo = torch.empty_like(h), torch.empty_like(x), torch.empty_like(h_bond)
        check(lib.pg_denoiser_forward(model.handle, self.handle, _ptr(h), _ptr(x), _ptr(h_bond), _ptr(phore_norm),
                                      _ptr(ho), _ptr(xo), _ptr(bo), _stream(self.device)), "pg_denoiser_forward")
        return ho, xo, bo

    @_on_device
    def phorediff_forward(self, model, h_node, pos, h_edge, time_step, h_phore_emb, pos_phore, phore_norm, out=None):
        h_node, pos, h_edge = _f32(h_node), _f32(pos), _f32(h_edge)
        assert h_node.shape == (self.Nl, 12) and pos.shape == (self.Nl, 3) and h_edge.shape == (self.Eb, 6)
        assert time_step.dtype == torch.int64 and time_step.shape == (self.G,) and time_step.is_contiguous()
        assert h_phore_emb.shape == (self.P, 128)
        h_phore_emb, pos_phore, phore_norm = _f32(h_phore_emb), _f32(pos_phore), _f32(phore_norm)   # keep converted copies alive
        if out is None:
            out = (_alloc((self.Nl, 12), torch.float32, self.device), _alloc((self.Nl, 3), torch.float32, self.device),
                   _alloc((self.Eb, 6), torch.float32, self.device))
        check(lib.pg_phorediff_forward(model.handle, self.handle, _ptr(h_node), _ptr(pos), _ptr(h_edge), _ptr(time_step),
                                       _ptr(h_phore_emb), _ptr(pos_phore), _ptr(phore_norm),
                                       _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream(self.device)), "pg_phorediff_forward")
        return out

    @_on_device
    def atom_count(self, model, h_phore_emb, h_phore, ex_col=12, min_atom=4, max_atom=78, intervals=False):
        """-> (count_l [G,1], count_u [G,1]) and, with intervals=True, the int32 bounds (lo [G], hi [G]) of sample_nodes."""
        h_phore_emb, h_phore = _f32(h_phore_emb), _f32(h_phore)
        assert h_phore_emb.shape == (self.P, 128) and h_phore.shape == (self.P, 18)
        cl, cu = _alloc((self.G, 1), torch.float32, self.device), _alloc((self.G, 1), torch.float32, self.device)
        lo = _alloc((self.G,), torch.int32, self.device) if intervals else None
        hi = _alloc((self.G,), torch.int32, self.device) if intervals else None
        check(lib.pg_atom_count(model.handle, self.handle, _ptr(h_phore_emb), _ptr(h_phore), ex_col, float(min_atom), float(max_atom),
                                _ptr(cl), _ptr(cu), _ptr(lo), _ptr(hi), _stream(self.device)), "pg_atom_count")
        return (cl, cu, lo, hi) if intervals else (cl, cu)

    # ---- per-molecule random streams ---------------------------------------------------------
    def molecule_streams(self, seed, graph_uid=None):
        """Per-molecule Philox addressing (see include/phoregen_b200.h, "Random-stream addressing"): molecule g draws from
        the key splitmix64(seed, uid_g) with its LOCAL row as counter, so its trajectory does not depend on the batch it
        is in or on the rank that runs it.  `graph_uid` [G] int: job-wide molecule ids (default: 0..G-1).
        -> dict(seed [G] int64 bit patterns, node_row0 [G] int64, edge_row0 [G] int64) on the plan's device."""
        uid = np.arange(self.G, dtype=np.uint64) if graph_uid is None else np.asarray(graph_uid).astype(np.uint64)
        assert uid.shape == (self.G,)
        with np.errstate(over="ignore"):
            z = (np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF) + (uid + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        na = self.num_atoms.astype(np.int64)
        node0 = np.concatenate([[0], np.cumsum(na)[:-1]])
        edge0 = np.concatenate([[0], np.cumsum(na * (na - 1))[:-1]])
        dev = self.device
        return dict(seed=torch.from_numpy(z.view(np.int64)).to(dev), node_row0=torch.from_numpy(node0).to(dev),
                    edge_row0=torch.from_numpy(edge0).to(dev))

    @staticmethod
    def _stream_args(streams, kind):
        if streams is None:
            return None, None
        return _ptr(streams["seed"]), _ptr(streams["node_row0" if kind == "node" else "edge_row0"])

    # ---- initial state (T4) --------------------------------------------------------------------
    @_on_device
    def sample_init(self, kind, log_prior, uniform=None, seed=0, streams=None):
        """transition.py:331-339: -> (onehot [rows,K] f32, cls [rows] int32, log_vt [rows,K] f32)."""
        K = 12 if kind == "node" else 6
        rows = self.Nl if kind == "node" else self.Eb
        log_prior = _f32(log_prior.to(self.device))
        assert log_prior.shape == (K,)
        onehot, cls = _alloc((rows, K), torch.float32, self.device), _alloc((rows,), torch.int32, self.device)
        log_vt = _alloc((rows, K), torch.float32, self.device)
        rg = self._lig_graph_ptr if kind == "node" else self._edge_graph_ptr
        gs, r0 = self._stream_args(streams, kind)
        check(lib.pg_sample_init(rows, K, _ptr(log_prior), ctypes.c_void_p(rg), _ptr(uniform), seed, 4 if kind == "node" else 5,
                                 _ptr(onehot), _ptr(cls), _ptr(log_vt), gs, r0, _stream(self.device)), "pg_sample_init")
        return onehot, cls, log_vt

    @_on_device
    def position_init(self, center=None, normal=None, seed=0, streams=None):
        """diffusion.py:406: standard normal minus the centre ([3] or one per graph [G,3])."""
        pos = _alloc((self.Nl, 3), torch.float32, self.device)
        gs, r0 = self._stream_args(streams, "node")
        check(lib.pg_position_init(self.Nl, ctypes.c_void_p(self._lig_graph_ptr), _ptr(normal), seed, 6, _ptr(center),
                                   int(center is not None and center.dim() == 2), _ptr(pos), gs, r0, _stream(self.device)), "pg_position_init")
        return pos

    # ---- transitions -----------------------------------------------------------------------
    @_on_device
    def categorical_step(self, model, kind, pred, log_vt, time_step, uniform=None, seed=0, step_counter=None,
                         onehot=None, cls=None, traj=None, streams=None):
        K = 12 if kind == "node" else 6
        rows = self.Nl if kind == "node" else self.Eb
        assert pred.shape == (rows, K) and log_vt.shape == (rows, K) and pred.is_contiguous() and log_vt.is_contiguous()
        qm = model.tables[f"{kind}_transition.q_mats"]
        tq = model.tables[f"{kind}_transition.transpopse_q_onestep_mats"]
        rg = self._lig_graph_ptr if kind == "node" else self._edge_graph_ptr
        if onehot is None:
            onehot = _alloc((rows, K), torch.float32, self.device)
        if cls is None:
            cls = _alloc((rows,), torch.int32, self.device)
        check(lib.pg_categorical_step(rows, K, _ptr(pred), _ptr(log_vt), _ptr(qm), _ptr(tq), _ptr(time_step),
                                      ctypes.c_void_p(rg), _ptr(uniform), seed, 1 if kind == "node" else 2,
                                      _ptr(step_counter), _ptr(onehot), _ptr(cls), _ptr(traj), *self._stream_args(streams, kind),
                                      _stream(self.device)), "pg_categorical_step")
        return onehot, cls

    @_on_device
    def position_step(self, model, x_t, x_recon, time_step, normal=None, energy_grad=None, seed=0, step_counter=None,
                      out=None, traj=None, center=None, streams=None):
        assert x_t.shape == (self.Nl, 3) and x_recon.shape == (self.Nl, 3)
        if out is None:
            out = torch.empty_like(x_t)
        t = model.tables
        check(lib.pg_position_step(self.Nl, _ptr(x_t), _ptr(x_recon), _ptr(energy_grad), _ptr(t["pos_transition.coef_x0"]),
                                   _ptr(t["pos_transition.coef_xt"]), _ptr(t["pos_transition.std"]), _ptr(time_step),
                                   ctypes.c_void_p(self._lig_graph_ptr), _ptr(normal), seed, 3, _ptr(step_counter),
                                   _ptr(out), _ptr(traj), _ptr(center), int(center is not None and center.dim() == 2),
                                   *self._stream_args(streams, "node"), _stream(self.device)), "pg_position_step")
        return out

    @_on_device
    def guidance_grad(self, pos, edge_cls, opts, phore_center, out=None, norm_graphs=0):
        """Sum of the drift gradients of every entry of `pos_guidance_opt` (diffusion.py:479-501 adds one gradient per
        list entry; unknown types contribute nothing there and are rejected here).  `phore_center`: [3] (one
        pharmacophore for the whole batch) or [G,3] (one per graph)."""
        if out is None:
            out = torch.empty_like(pos)
        per_graph = 1 if (phore_center is not None and phore_center.dim() == 2) else 0
        if per_graph:
            assert phore_center.shape == (self.G, 3)
        first = True
        for o in opts or []:
            if o["type"] == "atom_prox":
                flags, min_d, max_d = 1, float(o["min_d"]), float(o["max_d"])
            elif o["type"] == "center_prox":
                flags, min_d, max_d = 2, 0.0, 0.0
            else:
                raise NotImplementedError(f"pos_guidance_opt type {o['type']!r}")
            check(lib.pg_guidance_grad(self.handle, _ptr(pos), _ptr(edge_cls), flags | (0 if first else 4) | (8 * per_graph),
                                       min_d, max_d, _ptr(phore_center), _ptr(out), int(norm_graphs), _stream(self.device)), "pg_guidance_grad")
            first = False
        if first:
            out.zero_()
        return out
