"""GPU parity of the SURVEY 8f.3 row: window readout (gather), the reconstruction displacement kernel and FFTRecon
(algorithms/fftrecon.py) against the CPU oracle restatement (oracle/recon_oracle.py)."""
import numpy as np
import pytest

from oracle import pmesh_oracle as po
from oracle import recon_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def _pm(N, L, dtype):
    from nbodykit_b200.pmesh.pm import ParticleMesh
    from nbodykit_b200.comm import SelfComm
    return ParticleMesh(BoxSize=L, Nmesh=N, dtype=dtype, comm=SelfComm())


@pytest.mark.parametrize("resampler", ["nnb", "cic", "tsc", "pcs"])
@pytest.mark.parametrize("mesh_dtype,pos_dtype", [("f8", "f4"), ("f4", "f4"), ("f8", "f8")])
def test_readout_vs_oracle(cuda, resampler, mesh_dtype, pos_dtype):
    from nbodykit_b200.pmesh.pm import RealField
    N, L = [16, 24, 32], [32., 50., 10.]
    rng = np.random.RandomState(3)
    field = rng.standard_normal(N).astype(mesh_dtype)
    pos = (rng.uniform(-0.3, 1.3, size=(20000, 3)) * np.asarray(L)).astype(pos_dtype)   # includes out-of-box points
    pos[0] = 0.0
    pm = _pm(N, L, mesh_dtype)
    f = RealField(pm)
    f[...] = field
    got = f.readout(pos, resampler=resampler)
    want = ro.readout(field, pos, N, L, resampler)
    tol = 1e-12 if mesh_dtype == "f8" else 2e-6
    np.testing.assert_allclose(got, want, rtol=0, atol=tol * np.abs(field).max() * 8)
    # half-cell shifted transform, and the adjoint identity <paint(m), field> == <m, readout(field)>
    got2 = f.readout(pos, resampler=resampler, transform=pm.affine.shift(0.5))
    np.testing.assert_allclose(got2, ro.readout(field, pos, N, L, resampler, shift=0.5), rtol=0,
                               atol=tol * np.abs(field).max() * 8)
    m = rng.uniform(0.5, 1.5, size=len(pos))
    painted = po.paint(pos, m, N, L, resampler)
    assert abs((painted * field.astype("f8")).sum() - (m * want).sum()) < 1e-8 * np.abs(m * want).sum() + 1e-9


@pytest.mark.parametrize("dtype,tol", [("f8", 1e-13), ("f4", 1e-6)])
def test_recon_displacement_kernel(cuda, dtype, tol):
    import ctypes
    import torch
    from nbodykit_b200 import _lib
    from nbodykit_b200.pmesh.pm import ComplexField, RealField
    N, L = [16, 8, 32], [100., 50., 300.]
    rng = np.random.RandomState(5)
    real = rng.standard_normal(N).astype(dtype)
    pm = _pm(N, L, dtype)
    f = RealField(pm)
    f[...] = real
    ck = f.r2c()
    want_in = po.r2c(real.astype("f8"))
    los = np.array([0.6, 0.0, 0.8])
    out = ComplexField(pm)
    for axis in range(3):
        _lib.check(_lib.lib().nbk_recon_displacement(ctypes.c_void_p(ck.value.data_ptr()), ctypes.c_void_p(out.value.data_ptr()),
                                                     4 if dtype == "f4" else 8, _lib.iarr(N), _lib.darr(L), 0, 0, N[0], axis,
                                                     15.0, 1.7, 0.6, _lib.darr(los), None))
        torch.cuda.synchronize()
        want = ro.displacement_modes(want_in, N, L, axis, 15.0, 1.7, 0.6, los)
        got = out.numpy()
        assert np.abs(got - want).max() <= tol * np.abs(want).max() * 10
        assert got[0, 0, 0] == 0


@pytest.mark.parametrize("scheme,f,revert", [("LGS", 0.0, False), ("LRR", 0.4, False), ("LF2", 0.4, True)])
def test_fftrecon_vs_oracle(cuda, scheme, f, revert):
    """the three schemes on a small clustered catalogue: reconstructed mesh and the power spectrum of it"""
    from nbodykit_b200.lab import ArrayCatalog, FFTPower, FFTRecon
    N, L, R = 32, 400., 30.
    rng = np.random.RandomState(12)
    centres = rng.uniform(0, L, size=(300, 3))
    data = (centres[rng.randint(0, 300, size=20000)] + rng.standard_normal((20000, 3)) * 12.0) % L
    ran = rng.uniform(0, L, size=(60000, 3))
    dcat = ArrayCatalog({'Position': data}, BoxSize=L, Nmesh=N)
    rcat = ArrayCatalog({'Position': ran}, BoxSize=L, Nmesh=N)
    mesh = FFTRecon(data=dcat, ran=rcat, Nmesh=N, bias=1.5, f=f, los=[0, 0, 1], R=R, scheme=scheme,
                    revert_rsd_random=revert)
    got = mesh.compute(mode='real').numpy()
    want, s_d, s_r = ro.fftrecon(data, ran, N, L, bias=1.5, f=f, los=(0, 0, 1), R=R, scheme=scheme,
                                 revert_rsd_random=revert)
    # float32 shifted positions: a particle whose coordinate rounds differently moves weight between neighbouring cells
    # by ~1e-7 of a cell; fixed-point paint adds 1e-9.  Compare at 1e-4 of the field's dynamic range.
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    r = FFTPower(mesh, mode='1d')
    o = po.power_from_complex(po.r2c(want), None, N, L, mode='1d')
    assert np.array_equal(r.power['modes'], np.squeeze(o['modes']))
    np.testing.assert_allclose(r.power['power'].real, np.squeeze(o['power']).real, rtol=2e-4)
    assert mesh.attrs['scheme'] == scheme
