"""`.phore` files -> the pharmacophore tensors the sampler consumes (the step right before the
hot path).  Mirrors `PhoreData_New` (datasets/get_phore_data.py:12-105) and `AddPhoreNoise`
(datasets/transform.py:440-480, applied at sampling time through utils/training_utils.py:86-91) without PyG:
same 18-dimensional feature layout, same quirks, same random-number draws in the same order.

Feature row (data_name in zinc_300 / pdbbind): one-hot over
    MB HD AR PO HA HY NE CV1 CV2 CV3 CV4 XB EX  (13) | alpha (1) | has_norm one-hot (2) | exclusion-volume one-hot (2).
"""
import os

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from .testing import PhoreData

PHORE_TYPES = ("MB", "HD", "AR", "PO", "HA", "HY", "NE", "CV", "CR", "XB", "EX")                    # get_phore_data.py:8
PHORE_TYPES_CV = ("MB", "HD", "AR", "PO", "HA", "HY", "NE", "CV1", "CV2", "CV3", "CV4", "XB", "EX")  # get_phore_data.py:9


def read_phore_records(path, data_name="zinc_300"):
    """-> (type index list, alpha list, positions, has_norm list, normal end points) of the first block of the file.
    Line format (tab separated, get_phore_data.py:36-37):
        type alpha weight factor x y z has_norm norm_x norm_y norm_z label anchor_weight
    'CR' records are skipped, 'CV' is refined by the first character of its label, a line that does not parse is reported and
    skipped, '$$$$' ends the block (get_phore_data.py:30-52)."""
    if path is None or not os.path.exists(path):
        raise FileNotFoundError(f"The specified pharmacophore file (*.phore) is not found: `{path}`")
    names = PHORE_TYPES_CV if data_name in ("zinc_300", "pdbbind") else PHORE_TYPES
    index = {name: i for i, name in enumerate(names)}
    types, alphas, pos, has_norm, norm = [], [], [], [], []
    with open(path, "r") as f:
        f.readline()                                      # title
        for raw in f:
            rec = raw.strip()
            if rec == "$$$$":
                break
            try:
                (ptype, alpha, _weight, _factor, x, y, z, hn, nx, ny, nz, label, _anchor) = rec.split("\t")
                if ptype == "CR":
                    continue
                if ptype == "CV":
                    ptype += label[0]
                row = (index[ptype], float(alpha), [float(x), float(y), float(z)], int(hn), [float(nx), float(ny), float(nz)])
            except Exception as e:                        # the reference prints and goes on
                print(f"[E]: Failed to parse the line:\n {rec} | Message: {e}")
                continue
            types.append(row[0]); alphas.append(row[1]); pos.append(row[2]); has_norm.append(row[3]); norm.append(row[4])
    return types, alphas, pos, has_norm, norm, len(names)


def parse_phore_file(path, data_name="zinc_300", center="phore"):
    """PhoreData_New.get (get_phore_data.py:96-103) for one file: features, positions moved to the pharmacophore's centre
    of mass, unit normals, `center`.  Reference quirk kept: the *absolute* normal end point is normalised, not
    end point - position (get_phore_data.py:60-65)."""
    types, alphas, pos, has_norm, norm, n_types = read_phore_records(path, data_name)
    t = torch.nn.functional.one_hot(torch.tensor(types, dtype=torch.long), num_classes=n_types).float()
    ex = torch.nn.functional.one_hot(t[:, -1:].long().squeeze(-1), 2).float()
    hn = torch.nn.functional.one_hot(torch.tensor(has_norm, dtype=torch.long), 2).float()
    x = torch.cat((t, torch.tensor(alphas, dtype=torch.float).unsqueeze(-1), hn, ex), dim=-1)
    nrm = torch.tensor(norm, dtype=torch.float)
    length = nrm.norm(dim=-1, keepdim=True)
    nz = (length != 0).squeeze(-1)
    unit = torch.zeros_like(nrm)
    unit[nz] = nrm[nz] / length[nz]
    p = torch.tensor(pos, dtype=torch.float)
    com = p.mean(dim=0)
    if center == "phore":
        p = p - com
    elif center == "ligand":
        raise ValueError("center='ligand' needs a ligand; sampling uses center='phore' (sample_all.py:44)")
    name = os.path.splitext(os.path.basename(path))[0]
    return PhoreData(x, p, unit, center=com, name=name)


class AddPhoreNoise:
    """datasets/transform.py:440-480: Gaussian noise on the positions, and with probability 1/2 a rotation of every non-zero
    normal by an angle uniform in [0, `angle` degrees] about a random perpendicular axis.  Draw order (torch.randn_like;
    then per feature: numpy uniform for the angle, torch.rand for the coin, numpy uniform(2) for the axis) is the
    reference's, so equal global seeds give equal outputs."""

    def __init__(self, noise_std, angle):
        self.noise_std, self.angle = noise_std, angle

    @staticmethod
    def _perpendicular(v, epsilon=1e-12):
        a, b = np.random.uniform(0.1, 1, size=(2))
        if v[2] != 0:
            c = -(a * v[0] + b * v[1]) / v[2]
        else:
            assert not (v[0] == 0 and v[1] == 0)
            a, b, c = -v[1], v[0], 0
        axis = np.array([a, b, c])
        return axis / (np.linalg.norm(axis, axis=-1) + epsilon)

    def __call__(self, data):
        ph = data["phore"]
        ph["pos"] = ph["pos"] + torch.randn_like(ph["pos"]) * self.noise_std
        before = ph["norm"].clone()
        for i in range(before.size(0)):
            theta = np.random.uniform(0, np.pi / 180 * self.angle)
            if torch.all(before[i] == 0):
                continue
            if torch.rand(1) <= 0.5:
                v = before[i].numpy()
                turned = Rotation.from_rotvec(self._perpendicular(v) * theta).apply(v)
                ph["norm"][i] = torch.tensor(turned)
        return data


def collate_phores(items, copies=1):
    """Several pharmacophores in one sampler batch (the reference handles one per call, sample_all.py:69-94):
    -> dict(x, pos, norm, batch, center [G,3], names) for `TrajectorySampler(phore_batch=...)`, each item repeated `copies`
    times (int or one int per item)."""
    reps = [copies] * len(items) if isinstance(copies, int) else list(copies)
    xs, ps, ns, bs, cs, names, g = [], [], [], [], [], [], 0
    for it, r in zip(items, reps):
        ph = it["phore"]
        for _ in range(r):
            xs.append(ph["x"]); ps.append(ph["pos"]); ns.append(ph["norm"])
            bs.append(torch.full((ph["x"].shape[0],), g, dtype=torch.long))
            cs.append(it.center); names.append(it.name)
            g += 1
    return {"x": torch.cat(xs), "pos": torch.cat(ps), "norm": torch.cat(ns), "batch": torch.cat(bs),
            "center": torch.stack(cs), "names": names}
