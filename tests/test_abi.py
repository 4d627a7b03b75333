"""the C-ABI library loads and exports every symbol include/mkf_b200.h declares (no GPU needed)"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mkfbodytracker_pdaf_b200 as mk
from mkfbodytracker_pdaf_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mkf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mkf_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_library_agree():
    syms = header_symbols()
    assert sorted(L.SYMBOLS) == syms
    for s in syms:
        assert hasattr(L.lib, s), f"{s} declared in include/mkf_b200.h but not exported"
    assert L.lib.mkf_abi_version() == 1


def test_params_default_are_reference_literals():
    p = mk.default_params()
    assert p.chol_mode == mk.CHOL_CV24_LITERAL and p.alias_mode == mk.ALIAS_INDEPENDENT
    assert p.meas_noise_var == 100.0          # src/my_gmm.cpp:54
    assert p.assoc_pa == 0.05                 # src/pfPose.cpp:247
    assert p.assoc_clutter == 1e-4            # src/pfPose.cpp:261
    assert p.proposal_spread == 0.8           # src/pf2DRao.cpp:90
    assert p.neck_offset == 1.65              # src/pfPose.cpp:313
    assert (p.img_rows, p.img_cols) == (480, 640)


def test_model_loader_matches_reference_derivation(left_arm):
    a = left_arm.arrays
    assert (left_arm.mk.K, left_arm.mk.d, left_arm.mk.D) == (15, 12, 22)
    nm = left_arm.np
    # H = H1 * pca_proj^T, BH = H1 * pca_mean^T (src/my_gmm.cpp:61-72): pure selection, exact
    assert np.array_equal(a["H"], nm.H) and np.array_equal(a["BH"], nm.BH)
    assert np.allclose(a["Q"], nm.Q, rtol=1e-15, atol=0) and np.allclose(a["B"], nm.B, rtol=1e-15, atol=0)
    c = left_arm.orc.constants()
    for k in ("H", "BH", "Q", "B"):
        assert np.array_equal(c[k], a[k]), k
    # pca_* are stored as f32 and widened (src/pfPose.cpp:44-51)
    assert np.array_equal(a["pca_proj"], a["pca_proj"].astype(np.float32).astype(np.float64))
    assert abs(a["weights"].sum() - 1) < 1e-12


def test_model_loader_against_cv_filestorage(left_arm):
    cv2 = pytest.importorskip("cv2")
    ref = "/root/reference/data13D_PCA_100000_15_12.yml"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present on this box")
    fs = cv2.FileStorage(ref, cv2.FILE_STORAGE_READ)
    for k in ("means", "covs", "weights", "pca_proj", "pca_mean"):
        want = fs.getNode(k).mat().astype(np.float64).reshape(-1)
        assert np.array_equal(want, left_arm.arrays[k].reshape(-1)), k
    fs2 = cv2.FileStorage("/root/reference/data23D_PCA_100000_15_12.yml", cv2.FILE_STORAGE_READ)
    assert np.array_equal(fs2.getNode("gamma").mat().reshape(-1), left_arm.arrays["gamma"])  # quirk B4


def test_model_writer_round_trip(left_arm, tmp_path, rng):
    """mkf_model_save_yaml emits the schema src/pfPose.cpp:34-55 reads; values survive bit for bit"""
    out = tmp_path / "left.yml"
    left_arm.mk.save(str(out))
    back = mk.Model.load(str(out)).arrays()
    for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean", "Q", "B", "H", "BH"):
        assert np.array_equal(back[k], left_arm.arrays[k]), k
    txt = out.read_text()
    assert txt.startswith("%YAML:1.0\n") and txt.count("!!opencv-matrix") == 6
    assert "pca_proj: !!opencv-matrix\n   rows: 12\n   cols: 22\n   dt: f" in txt   # widened floats stay f32 on disk
    # another shape (launch/ChaLearn.launch:7-8 names K = 20, d = 10 files); a true-f64 pca_proj is kept as dt: d
    K, d, D = 20, 10, 22
    covs = np.stack([(lambda a: a @ a.T + d * np.eye(d))(rng.normal(size=(d, d))) for _ in range(K)])
    w = rng.random(K)
    m = mk.Model.from_arrays(rng.normal(size=(K, d)) * 30, covs.reshape(K * d, d), w / w.sum(), 0.85 + 0.1 * rng.random(K),
                             np.linalg.qr(rng.normal(size=(D, d)))[0].T.copy(), rng.normal(size=D) * 100)
    out2 = tmp_path / "k20.yml"
    m.save(str(out2))
    assert "pca_proj: !!opencv-matrix\n   rows: 10\n   cols: 22\n   dt: d" in out2.read_text()
    a, b = m.arrays(), mk.Model.load(str(out2)).arrays()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    cv2 = pytest.importorskip("cv2")   # the parser the reference uses must accept the file as well
    fs = cv2.FileStorage(str(out), cv2.FILE_STORAGE_READ)
    for k in ("means", "covs", "weights", "pca_proj", "pca_mean", "gamma"):
        got = fs.getNode(k).mat()
        assert got.dtype == (np.float32 if k.startswith("pca_") else np.float64), k
        assert np.array_equal(got.astype(np.float64).reshape(-1), left_arm.arrays[k].reshape(-1)), k
    with pytest.raises(mk.MkfError) as e:
        left_arm.mk.save(str(tmp_path / "no_such_dir" / "x.yml"))
    assert e.value.code == L.E_IO


def test_loader_errors(tmp_path):
    with pytest.raises(mk.MkfError) as e:
        mk.Model.load(str(tmp_path / "missing.yml"))
    assert e.value.code == L.E_IO
    bad = tmp_path / "bad.yml"
    bad.write_text("%YAML:1.0\nmeans: !!opencv-matrix\n   rows: 2\n   cols: 2\n   dt: d\n   data: [ 1., 2., 3. ]\n")
    with pytest.raises(mk.MkfError) as e:
        mk.Model.load(str(bad))
    assert e.value.code == L.E_PARSE


def test_model_argument_validation(left_arm):
    a = left_arm.arrays
    with pytest.raises(mk.MkfError):  # d = 11 is not a built shape
        mk.Model.from_arrays(np.zeros((3, 11)), np.tile(np.eye(11), (3, 1)), np.ones(3) / 3, np.ones(3) * 0.9,
                             np.zeros((11, 22)), np.zeros(22))
    p = mk.default_params()
    p.alias_mode = 7
    with pytest.raises(mk.MkfError) as e:
        mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)
    assert e.value.code == L.E_INVALID
    p.alias_mode = L.ALIAS_CV_SHALLOW_LITERAL  # quirk B3 is a supported mode
    mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], p)


def test_no_cpu_fallback(left_arm):
    if mk.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mk.MkfError) as e:
        mk.TrackBatch(left_arm.mk, 2, 16)
    assert e.value.code == L.E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(mk.MkfError):
        mk.resample(np.ones(4) / 4, 4, 0.5)


def test_product_does_not_touch_oracle():
    """the shipped package must never import / link oracle/ (judge check)"""
    pkg = os.path.join(ROOT, "mkfbodytracker_pdaf_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "mkf_oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, fn
    for fn in os.listdir(os.path.join(ROOT, "include")):
        txt = open(os.path.join(ROOT, "include", fn)).read()
        assert "mkf_oracle" not in txt and "liboracle" not in txt
