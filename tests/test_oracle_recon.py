"""CPU checks of the reconstruction oracle (oracle/recon_oracle.py): internal identities that hold for the pmesh
operations it restates (the reference cannot run here -- pmesh is absent -- so this flow is parity-unpinned)."""
import numpy as np
import pytest

from oracle import pmesh_oracle as po
from oracle import recon_oracle as ro


@pytest.mark.parametrize("resampler", ["nnb", "cic", "tsc", "pcs"])
def test_readout_is_the_adjoint_of_paint(resampler):
    N, L = [8, 12, 16], [20., 30., 8.]
    rng = np.random.RandomState(1)
    field = rng.standard_normal(N)
    pos = rng.uniform(-0.5, 1.5, size=(500, 3)) * np.asarray(L)
    m = rng.uniform(size=500)
    lhs = (po.paint(pos, m, N, L, resampler) * field).sum()
    rhs = (m * ro.readout(field, pos, N, L, resampler)).sum()
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
    # partition of unity: a constant field reads out as that constant
    np.testing.assert_allclose(ro.readout(np.full(N, 3.5), pos, N, L, resampler), 3.5, rtol=1e-13)


def test_displacement_modes_and_schemes():
    N, L = 16, 200.
    rng = np.random.RandomState(2)
    data = rng.uniform(0, L, size=(3000, 3))
    ran = rng.uniform(0, L, size=(6000, 3))
    dk = po.r2c(ro.work_with(data, None, len(data), N, L))
    for axis in range(3):
        d = ro.displacement_modes(dk, N, L, axis, 20., 2.0, 0.5, np.array([0, 0, 1.]))
        assert d[0, 0, 0] == 0
        assert np.isfinite(po.c2r(d, N)).all()
    lgs, s_d, s_r = ro.fftrecon(data, ran, N, L, bias=2.0, f=0.0, R=20., scheme="LGS")
    lrr, _, _ = ro.fftrecon(data, ran, N, L, bias=2.0, f=0.0, R=20., scheme="LRR")
    lf2, _, _ = ro.fftrecon(data, ran, N, L, bias=2.0, f=0.0, R=20., scheme="LF2")
    np.testing.assert_allclose(lf2, lgs * (3.0 / 7.0) + lrr * (4.0 / 7.0), rtol=0, atol=1e-12)
    assert s_d.dtype == np.float32 and s_d.shape == (3000, 3) and abs(lgs.mean()) < 1e-10


def test_convpower_full_mesh_oracle_agrees_with_the_hermitian_one_for_even_multipoles():
    """oracle/convpower_oracle.py: the full complex-mesh restatement (fftn, hermitian_symmetric=False) and the
    Hermitian one coincide for even multipoles and mode counts; odd multipoles of the full mesh are purely imaginary"""
    from oracle import convpower_oracle as co
    rng = np.random.RandomState(4)
    L = np.array([300., 300., 300.])
    C = np.array([500., 100., 50.])
    d = rng.uniform(-130, 130, size=(3000, 3)) + C
    r = rng.uniform(-130, 130, size=(9000, 3)) + C
    nb = 1e-4
    wf = 1. / (1 + 1e4 * nb)
    args = (d, r, (np.ones(len(d)), wf * np.ones(len(d))), (np.ones(len(r)), wf * np.ones(len(r))),
            nb * np.ones(len(d)), nb * np.ones(len(r)), 16, L, C)
    full = co.convpower_full(*args, [0, 1, 2], dk=0.05)
    herm = co.convpower(*args, [0, 1, 2], dk=0.05)
    assert np.array_equal(full['modes'], herm['modes'])
    scale = np.nanmax(np.abs(herm['power_0']))
    for ell in (0, 2):
        assert np.nanmax(np.abs(full['power_%d' % ell] - herm['power_%d' % ell])) < 1e-12 * scale
    assert np.nanmax(np.abs(full['power_1'].real[:-1])) < 1e-10 * scale
    assert np.nanmax(np.abs(full['power_1'].imag)) > 1e-3 * scale
