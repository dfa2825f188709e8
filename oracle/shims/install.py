"""TEST INFRASTRUCTURE ONLY — pure-torch stand-ins for the PyG stack so that the
UNMODIFIED reference (`/root/reference/models/*.py`) imports in this container.

Used only by `oracle/make_golden.py` (fixture generation, run where
/root/reference exists) and by `tests/test_oracle_vs_reference.py`.  Nothing
under `phoregen_b200/` may import this module.

Third-party packages stood in for (pinned in reference `phoregen_env.yml:314-317`):
  torch-cluster 1.6.0   -> knn / knn_graph       (call sites uni_denoiser.py:355, common.py:245,301)
  torch-scatter 2.0.9   -> scatter, scatter_sum, scatter_softmax
                           (uni_denoiser.py:62,66,158,162,204,208; diffusion.py:150,157; common.py:287,295,303)
  torch-sparse 0.6.15   -> SparseTensor           (uni_denoiser.py:105-121)
  torch-geometric 2.1.0 -> HeteroData / Batch / Dataset / remove_self_loops
rdkit / openbabel / easydict are only needed for `import` to succeed.

kNN semantics restated from torch_cluster 1.6.0's CUDA kernel (`knn_cuda.cu`):
per query, scan the candidates of the same graph in index order, squared
distance accumulated in fp32 over d = 0,1,2, keep the best k with strict '<'
insertion (ties -> lower index first); knn_graph(loop=False) asks for k+1 and
drops the self edge.  Here: stable argsort of ((xi-xj)^2).sum(-1).
"""
import sys
import types
from unittest import mock

import torch


# ----------------------------------------------------------------- torch_scatter
def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src), dim


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    idx, dim = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    res.scatter_add_(dim, idx, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(shape, dtype=src.dtype, device=src.device)
        cnt.scatter_add_(dim, idx, torch.ones_like(src))
        return res / cnt.clamp(min=1)
    raise NotImplementedError(reduce)


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim=dim, dim_size=dim_size, reduce="sum")


def scatter_softmax(src, index, dim=-1, dim_size=None):
    idx, dim = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    mx = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    mx.scatter_reduce_(dim, idx, src, reduce="amax", include_self=True)
    rec = src - mx.gather(dim, idx)
    ex = rec.exp()
    sm = torch.zeros(shape, dtype=src.dtype, device=src.device)
    sm.scatter_add_(dim, idx, ex)
    return ex / sm.gather(dim, idx)


# ----------------------------------------------------------------- torch_cluster / PyG nn
def knn(x, y, k, batch_x=None, batch_y=None):
    """For each y: the k nearest x of the same graph.  Returns [2, E] (y_idx, x_idx)."""
    if batch_x is None:
        batch_x = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(y.size(0), dtype=torch.long, device=y.device)
    rows, cols = [], []
    for b in torch.unique(batch_y).tolist():
        iy = (batch_y == b).nonzero()[:, 0]
        ix = (batch_x == b).nonzero()[:, 0]
        if ix.numel() == 0:
            continue
        d = ((y[iy][:, None, :] - x[ix][None, :, :]) ** 2).sum(-1)
        order = torch.sort(d, dim=1, stable=True).indices[:, :k]
        kk = order.size(1)
        rows.append(iy[:, None].expand(-1, kk).reshape(-1))
        cols.append(ix[order].reshape(-1))
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], 0)


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **kw):
    ei = knn(x, x, k if loop else k + 1, batch, batch)
    if flow == "source_to_target":
        row, col = ei[1], ei[0]
    else:
        row, col = ei[0], ei[1]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)


def _not_needed(*a, **k):
    raise NotImplementedError("shim: not on the hot path")


def remove_self_loops(edge_index, edge_attr=None):
    m = edge_index[0] != edge_index[1]
    return edge_index[:, m], (None if edge_attr is None else edge_attr[m])


# ----------------------------------------------------------------- torch_sparse
class _Storage:
    def __init__(self, row, col, value):
        self._row, self._col, self._value = row, col, value

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value


class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None, _sorted=False):
        if not _sorted:
            key = row * (int(sparse_sizes[1]) + 1) + col
            perm = torch.sort(key, stable=True).indices
            row, col = row[perm], col[perm]
            value = None if value is None else value[perm]
        self.storage = _Storage(row, col, value)
        self.sizes = tuple(int(s) for s in sparse_sizes)

    def __getitem__(self, idx):
        row, col, val = self.storage._row, self.storage._col, self.storage._value
        n_rows = self.sizes[0]
        counts = torch.bincount(row, minlength=n_rows)
        ptr = torch.zeros(n_rows + 1, dtype=torch.long, device=row.device)
        ptr[1:] = torch.cumsum(counts, 0)
        cnt = counts[idx]
        new_row = torch.repeat_interleave(torch.arange(idx.numel(), device=row.device), cnt)
        start = ptr[idx]
        off = torch.arange(int(cnt.sum()), device=row.device) - torch.repeat_interleave(
            torch.cumsum(cnt, 0) - cnt, cnt)
        src = torch.repeat_interleave(start, cnt) + off
        return SparseTensor(new_row, col[src], None if val is None else val[src],
                            sparse_sizes=(idx.numel(), self.sizes[1]), _sorted=True)

    def set_value(self, value, layout=None):
        return SparseTensor(self.storage._row, self.storage._col, value,
                            sparse_sizes=self.sizes, _sorted=True)

    def sum(self, dim):
        assert dim == 1
        v = self.storage._value
        if v is None:
            v = torch.ones(self.storage._row.numel(), device=self.storage._row.device)
        out = torch.zeros(self.sizes[0], dtype=v.dtype, device=v.device)
        out.index_add_(0, self.storage._row, v)
        return out


# ----------------------------------------------------------------- torch_geometric.data
class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @property
    def num_nodes(self):
        for key in ("x", "pos"):
            if key in self:
                return self[key].size(0)
        return 0


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_stores", {})
        object.__setattr__(self, "_attrs", {})

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2:
            key = (key[0], "to", key[1]) if not any(
                isinstance(k, tuple) and k[0] == key[0] and k[-1] == key[1] for k in self._stores
            ) else next(k for k in self._stores if isinstance(k, tuple) and k[0] == key[0] and k[-1] == key[1])
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_attrs")[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self._attrs[k] = v

    def clone(self):
        out = HeteroData()
        for k, st in self._stores.items():
            ns = out[k]
            for kk, v in st.items():
                ns[kk] = v.clone() if torch.is_tensor(v) else v
        for k, v in self._attrs.items():
            out._attrs[k] = v.clone() if torch.is_tensor(v) else v
        return out

    def to(self, device):
        for st in self._stores.values():
            for kk, v in list(st.items()):
                if torch.is_tensor(v):
                    st[kk] = v.to(device)
        for k, v in list(self._attrs.items()):
            if torch.is_tensor(v):
                self._attrs[k] = v.to(device)
        return self


class Batch(HeteroData):
    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        out = cls()
        keys = data_list[0]._stores.keys()
        for k in keys:
            st = out[k]
            fields = data_list[0]._stores[k].keys()
            for f in fields:
                vals = [d._stores[k][f] for d in data_list]
                if torch.is_tensor(vals[0]) and vals[0].dim() >= 1:
                    st[f] = torch.cat(vals, 0)
            if "x" in fields or "pos" in fields:
                ref = "x" if "x" in fields else "pos"
                sizes = [d._stores[k][ref].size(0) for d in data_list]
                st["batch"] = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(sizes))
                ptr = torch.zeros(len(sizes) + 1, dtype=torch.long)
                ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
                st["ptr"] = ptr
        out._attrs["num_graphs"] = len(data_list)
        return out


class Dataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.transform = transform

    def __len__(self):
        return self.len()

    def __getitem__(self, idx):
        d = self.get(idx)
        return d if self.transform is None else self.transform(d)


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, list):
            v = [EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v]
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def install(reference_root="/root/reference"):
    """Register the stand-in modules and put the reference on sys.path."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("torch_scatter", scatter=scatter, scatter_sum=scatter_sum, scatter_softmax=scatter_softmax)
    mod("torch_sparse", SparseTensor=SparseTensor)
    mod("torch_cluster", knn=knn, knn_graph=knn_graph)
    tg = mod("torch_geometric")
    tg.nn = mod("torch_geometric.nn", knn_graph=knn_graph, knn=knn, radius_graph=_not_needed, radius=_not_needed)
    def k_hop_subgraph(*args, **kwargs):      # only MaskByPhore (training-time masking, outside the sampling path) calls it
        raise NotImplementedError("torch_geometric.utils.k_hop_subgraph is not provided by the stand-ins")
    tg.utils = mod("torch_geometric.utils", remove_self_loops=remove_self_loops, k_hop_subgraph=k_hop_subgraph)
    tg.data = mod("torch_geometric.data", HeteroData=HeteroData, Batch=Batch, Dataset=Dataset)
    mod("easydict", EasyDict=EasyDict)
    for name in ("rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.Geometry", "rdkit.RDLogger",
                 "rdkit.Chem.rdchem", "rdkit.Chem.rdMolTransforms", "openbabel", "openbabel.openbabel",
                 "rdkit.Chem.Descriptors", "rdkit.Chem.rdMolDescriptors", "rdkit.Chem.Draw"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
