"""Pin the oracle restatement directly against the reference's code (only where /root/reference exists)."""
import numpy as np
import pytest

from oracle import refload, pmesh_oracle as po

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ns():
    return refload.load()


@pytest.mark.parametrize("N,L,cd,coord,Nmu,poles,los", [
    ([16, 16, 16], [64.] * 3, "c16", "f4", 5, [0, 2, 4], [0, 0, 1]),
    ([12, 8, 10], [100., 50., 70.], "c8", "f4", 3, [1, 2], [0, 1, 0]),
    ([16, 16, 16], [100.] * 3, "c16", "f8", 4, [3], [0.6, 0.0, 0.8]),
    ([8, 8, 8], [1.] * 3, "c16", "f4", 1, [], [0, 0, 1]),
])
def test_project_to_basis(ns, N, L, cd, coord, Nmu, poles, los):
    rng = np.random.RandomState(3)
    shape = (N[0], N[1], N[2] // 2 + 1)
    y = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cd)
    x = po.k_coords(N, L, coord)
    dk = 2 * np.pi / min(L)
    kedges = np.arange(0., np.pi * min(N) / max(L) + dk / 2, dk)
    muedges = np.linspace(-1, 1, Nmu + 1)
    ref, pref = ns.project_to_basis(refload.RefComplexField(y, x), [kedges, muedges], los=los, poles=poles)
    got, pgot = po.project_to_basis(y, x, [kedges, muedges], los=los, poles=poles)
    assert np.array_equal(ref[3], got[3])
    tol = 1e-12 if cd == "c16" else 1e-6
    for a, b in zip(ref[:3], got[:3]):
        np.testing.assert_allclose(b, a, rtol=tol, atol=tol, equal_nan=True)
    if poles:
        assert np.array_equal(pref[2], pgot[2])
        np.testing.assert_allclose(pgot[1], pref[1], rtol=tol, atol=tol, equal_nan=True)


def test_compensation_functions(ns):
    N, L = [8, 16, 12], [10., 20., 30.]
    rng = np.random.RandomState(4)
    v = rng.standard_normal((8, 16, 7)) + 0j
    for coord in ["f4", "f8"]:
        w = po.k_coords(N, L, coord, kind="circular")
        for interlaced in (True, False):
            for res in ("cic", "tsc", "pcs"):
                func = ns.get_compensation(interlaced, res)[0][1]
                assert func.__name__ == po.COMPENSATION[(interlaced, res)]
                np.testing.assert_array_equal(po.compensate(func.__name__, w, v.copy()), func(w, v.copy()))


def test_mpirng(ns):
    ref = ns.MPIRandomState(ns.FakeComm(), seed=7, size=123456)
    mine = po.SerialMPIRandomState(7, 123456)
    np.testing.assert_array_equal(ref.uniform(itemshape=(3,)), mine.uniform(itemshape=(3,)))
    np.testing.assert_array_equal(ref.normal(), mine.normal())


def test_product_mpirng_matches_reference(ns):
    from nbodykit_b200.mpirng import MPIRandomState
    from nbodykit_b200.comm import SelfComm
    ref = ns.MPIRandomState(ns.FakeComm(), seed=9, size=250001)
    mine = MPIRandomState(SelfComm(), seed=9, size=250001)
    np.testing.assert_array_equal(ref.uniform(itemshape=(3,)), mine.uniform(itemshape=(3,)))
    lam = np.linspace(0.5, 3, 250001)
    np.testing.assert_array_equal(ref.poisson(lam=lam), mine.poisson(lam=lam))
    np.testing.assert_array_equal(ref.normal(loc=1., scale=3.), mine.normal(loc=1., scale=3.))
