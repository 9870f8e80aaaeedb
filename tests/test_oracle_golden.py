"""The oracle restatement against golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py) and against the reference fixture nbodykit/tests/data/dataset_2d.json."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import pmesh_oracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _field(N, seed, cdtype):
    rng = np.random.RandomState(seed)
    shape = (N[0], N[1], N[2] // 2 + 1)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cdtype)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "project_to_basis_*.npz"))))
def test_project_to_basis_golden(path):
    g = np.load(path)
    N, L = g["N"].tolist(), g["L"].tolist()
    y = _field(N, int(g["seed"]), str(g["cdtype"]))
    x = po.k_coords(N, L, str(g["coord"]))
    los = [float(v) if v != int(v) else int(v) for v in g["los"]]
    res, pres = po.project_to_basis(y, x, [g["kedges"], g["muedges"]], los=los, poles=g["poles"].tolist())
    assert np.array_equal(res[3], g["N2d"])                      # mode counts: exact
    tol = 1e-12 if str(g["cdtype"]) == "c16" else 1e-6           # reference sums c8 fields in c8
    np.testing.assert_allclose(res[0], g["xmean"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(res[1], g["mumean"], rtol=1e-12, atol=1e-15, equal_nan=True)
    np.testing.assert_allclose(res[2], g["y2d"], rtol=tol, atol=tol, equal_nan=True)
    if len(g["poles"]):
        assert np.array_equal(pres[2], g["pole_N"])
        np.testing.assert_allclose(pres[0], g["pole_k"], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(pres[1], g["pole_y"], rtol=tol, atol=tol, equal_nan=True)


def test_compensate_golden():
    g = np.load(os.path.join(GOLD, "compensate.npz"))
    N, L = g["N"].tolist(), g["L"].tolist()
    w = po.k_coords(N, L, "f4", kind="circular")
    v = _field(N, int(g["seed"]), "c16")
    for name in po.COMPENSATION.values():
        np.testing.assert_array_equal(po.compensate(name, w, v.copy()), g[name])


def test_mpirng_golden():
    g = np.load(os.path.join(GOLD, "mpirng.npz"))
    rng = po.SerialMPIRandomState(42, 250000)
    u1 = rng.uniform(itemshape=(3,))
    u2 = rng.uniform(itemshape=(3,))
    np.testing.assert_array_equal(u1[:5], g["uniform_first"])
    np.testing.assert_array_equal(u1[[0, 99999, 100000, 199999, 200000, 249999]], g["uniform_rows"])
    np.testing.assert_array_equal(np.array([u1.sum(), u2.sum()]), g["uniform_sum"])
    np.testing.assert_array_equal(rng.normal(loc=1.0, scale=2.0)[[0, 100000, 249999]], g["normal_rows"])
    assert np.random.RandomState(42).poisson(1e5) == g["N_uniformcatalog"][0] == 99886
    assert np.random.RandomState(42).poisson(100) == g["N_uniformcatalog"][1] == 96   # fftpower.ipynb:313


def test_uniform_catalog_first_rows():
    pos, vel = po.uniform_catalog(100, 1.0, 42)
    assert len(pos) == 96
    np.testing.assert_allclose(pos[0], [0.45470105, 0.83263203, 0.06905134], atol=5e-9)


def test_shell_counts_fixture():
    """float32 coordinate arithmetic reproduces the reference fixture's per-shell mode counts exactly;
    float64 arithmetic does not (lattice modes that sit on a bin edge) -- SURVEY B.5 / C.4"""
    gold = json.load(open(os.path.join(GOLD, "dataset_2d_modes.json")))
    N, L = gold["Nmesh"], gold["BoxSize"]
    dk = 2 * np.pi / L
    kedges = np.arange(0., np.pi * N / L + dk / 2, dk)
    c = np.ones((N, N, N // 2 + 1), dtype="c8")
    res, _ = po.project_to_basis(c, po.k_coords(N, L, "f4"), [kedges, np.linspace(-1, 1, 2)])
    assert res[3][:, 0].tolist() == gold["modes_k"]
    res8, _ = po.project_to_basis(c, po.k_coords(N, L, "f8"), [kedges, np.linspace(-1, 1, 2)])
    assert res8[3][:, 0].sum() == sum(gold["modes_k"]) and res8[3][:, 0].tolist() != gold["modes_k"]


def test_paint_conserves_mass_and_is_linear():
    """source/mesh/tests/test_species.py:50-75 style identities on the oracle paint"""
    rng = np.random.RandomState(0)
    pos = rng.uniform(-10, 50, size=(5000, 3))
    w = rng.uniform(size=5000)
    for res in ["nnb", "cic", "tsc", "pcs"]:
        a = po.paint(pos, w, 16, 32., res)
        np.testing.assert_allclose(a.sum(), w.sum(), rtol=1e-12)
        b = po.paint(pos[:2000], w[:2000], 16, 32., res) + po.paint(pos[2000:], w[2000:], 16, 32., res)
        np.testing.assert_allclose(a, b, atol=1e-12)
        # slab painting with ghost dropping reassembles the full mesh
        parts = [po.paint(pos, w, 16, 32., res, x_start=s, x_n=4) for s in range(0, 16, 4)]
        np.testing.assert_allclose(np.concatenate(parts, axis=0), a, atol=1e-12)


def test_node_particle_deposits_fully():
    """a particle exactly on a mesh node deposits all its mass there"""
    pos = np.array([[3.0, 5.0, 7.0]]) * (32. / 16)
    for res in ["nnb", "cic"]:
        m = po.paint(pos, None, 16, 32., res)
        assert m[3, 5, 7] == 1.0 and m.sum() == 1.0


def test_r2c_normalisation_and_roundtrip():
    rng = np.random.RandomState(1)
    real = rng.standard_normal((8, 16, 4))
    c = po.r2c(real)
    np.testing.assert_allclose(c[0, 0, 0].real, real.mean(), rtol=1e-12)     # forward carries 1/N^3
    np.testing.assert_allclose(po.c2r(c, real.shape), real, atol=1e-12)      # backward unnormalised
