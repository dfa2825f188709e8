"""CPU tests of the `.phore` reader and the sampling-time noise transform (phoregen_b200/phore_io.py) against the
unmodified reference classes (datasets/get_phore_data.py, datasets/transform.py) where /root/reference is mounted, and
against hand-computed values otherwise."""
import os

import numpy as np
import pytest
import torch

from phoregen_b200 import phore_io

HAVE_REF = os.path.isdir("/root/reference/models")

ROWS = [  # type alpha weight factor x y z has_norm nx ny nz label anchor_weight
    ("HD", 0.7, 1.0, 1.0, 1.0, 2.0, 3.0, 1, 2.0, 2.5, 3.5, "0", 1.0),
    ("CR", 0.7, 1.0, 1.0, 9.0, 9.0, 9.0, 0, 0.0, 0.0, 0.0, "0", 1.0),       # skipped
    ("CV", 0.5, 1.0, 1.0, -1.0, 0.5, 0.0, 0, 0.0, 0.0, 0.0, "3", 1.0),      # -> CV3
    ("AR", 0.9, 1.0, 1.0, 0.0, -2.0, 1.5, 1, 0.0, -2.0, 3.0, "0", 1.0),
    ("EX", 0.837, 0.5, 1.0, 4.0, 4.0, -4.0, 0, 0.0, 0.0, 0.0, "0", 1.0),    # exclusion volume
    ("HA", 0.7, 1.0, 1.0, -3.0, 1.0, 1.0, 1, -3.0, 1.0, 2.0, "0", 1.0),
]


def _write(path):
    with open(path, "w") as f:
        f.write("synthetic_pharmacophore\n")
        for r in ROWS[:3]:
            f.write("\t".join(str(v) for v in r) + "\n")
        f.write("this line does not parse\n")
        for r in ROWS[3:]:
            f.write("\t".join(str(v) for v in r) + "\n")
        f.write("$$$$\n")
        f.write("\t".join(str(v) for v in ROWS[0]) + "\n")                  # second block: ignored
    return path


def test_parse_matches_hand_computed_layout(tmp_path):
    d = phore_io.parse_phore_file(_write(str(tmp_path / "m1.phore")))
    ph = d["phore"]
    assert d.name == "m1" and ph["x"].shape == (5, 18) and ph["pos"].shape == (5, 3)
    types = ph["x"][:, :13].argmax(-1).tolist()
    assert types == [1, 9, 2, 12, 4]                                        # HD, CV3, AR, EX, HA
    assert torch.allclose(ph["x"][:, 13], torch.tensor([0.7, 0.5, 0.9, 0.837, 0.7]))
    assert ph["x"][:, 14:16].argmax(-1).tolist() == [1, 0, 1, 0, 1]         # has_norm one-hot
    assert ph["x"][:, 16:18].argmax(-1).tolist() == [0, 0, 0, 1, 0]         # exclusion-volume one-hot
    raw = torch.tensor([r[4:7] for r in ROWS if r[0] != "CR"], dtype=torch.float)
    assert torch.allclose(d.center, raw.mean(0)) and torch.allclose(ph["pos"], raw - raw.mean(0))
    n0 = torch.tensor([2.0, 2.5, 3.5])
    assert torch.allclose(ph["norm"][0], n0 / n0.norm()) and torch.all(ph["norm"][1] == 0)    # absolute end point normalised
    with pytest.raises(FileNotFoundError):
        phore_io.parse_phore_file(str(tmp_path / "missing.phore"))


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_parse_and_noise_match_the_unmodified_reference(tmp_path):
    from oracle.shims.install import install
    install()
    from datasets.get_phore_data import PhoreData_New
    from datasets.transform import AddPhoreNoise
    path = _write(str(tmp_path / "m2.phore"))
    want = PhoreData_New([path], center="phore", data_name="zinc_300").get(0)
    got = phore_io.parse_phore_file(path)
    for key in ("x", "pos", "norm"):
        assert torch.equal(got["phore"][key], getattr(want["phore"], key)), key
    assert torch.equal(got.center, want.center) and got.name == want.name
    changed = 0
    for seed in (0, 1, 2, 3):
        a = phore_io.parse_phore_file(path)
        b = PhoreData_New([path], center="phore", data_name="zinc_300").get(0)
        torch.manual_seed(seed); np.random.seed(seed)
        a = phore_io.AddPhoreNoise(0.1, 5.0)(a)
        torch.manual_seed(seed); np.random.seed(seed)
        b = AddPhoreNoise(noise_std=0.1, angle=5.0)(b)
        assert torch.equal(a["phore"]["pos"], b["phore"].pos) and torch.equal(a["phore"]["norm"], b["phore"].norm)
        changed += int(not torch.equal(a["phore"]["norm"], got["phore"]["norm"]))
    assert changed > 0          # the rotation branch was exercised


def test_collate_phores(tmp_path):
    d = phore_io.parse_phore_file(_write(str(tmp_path / "m3.phore")))
    b = phore_io.collate_phores([d, d], copies=[2, 1])
    assert b["x"].shape == (15, 18) and b["batch"].tolist() == [0] * 5 + [1] * 5 + [2] * 5
    assert b["center"].shape == (3, 3) and b["names"] == ["m3"] * 3


# ---------------------------------------------------------------- the reference's own sampling inputs (configs[0])
PHORE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phores")


def shipped_phore_files():
    """The 10 files of reference data/phores_for_sampling (file_index.json order), kept as input fixtures."""
    import json
    index = json.load(open(os.path.join(PHORE_DIR, "file_index.json")))
    return [os.path.join(PHORE_DIR, os.path.basename(p)) for p in index]


def test_shipped_phore_files_parse_to_the_surveyed_shapes():
    files = shipped_phore_files()
    assert len(files) == 10
    sizes = []
    for f in files:
        d = phore_io.parse_phore_file(f)
        ph = d["phore"]
        n, n_ex = ph["x"].shape[0], int((ph["x"][:, 12] == 1).sum())
        sizes.append((n, n_ex))
        assert ph["x"].shape == (n, 18) and torch.all(ph["x"][:, :13].sum(-1) == 1) and torch.all(ph["x"][:, 16:18].sum(-1) == 1)
        assert torch.allclose(ph["pos"].mean(0), torch.zeros(3), atol=1e-4) and 3 <= n - n_ex <= 7
    assert min(s[0] for s in sizes) == 44 and max(s[0] for s in sizes) == 99          # SURVEY.md §8(d) config[0]
    assert min(s[1] for s in sizes) == 40 and max(s[1] for s in sizes) == 94


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_shipped_phore_files_match_the_unmodified_reference_reader():
    from oracle.shims.install import install
    install()
    from datasets.get_phore_data import PhoreData_New
    files = shipped_phore_files()
    ref_files = [os.path.join("/root/reference/data/phores_for_sampling", os.path.basename(f)) for f in files]
    ds = PhoreData_New(ref_files, center="phore", data_name="zinc_300")
    for i, f in enumerate(files):
        assert open(f, "rb").read() == open(ref_files[i], "rb").read()               # the fixture IS the reference's input
        want, got = ds.get(i), phore_io.parse_phore_file(f)
        for key in ("x", "pos", "norm"):
            assert torch.equal(got["phore"][key], getattr(want["phore"], key)), (f, key)
        assert torch.equal(got.center, want.center) and got.name == want.name
