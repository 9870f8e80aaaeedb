"""MeshSource API on the GPU: FieldMesh / ArrayMesh sources, user callbacks through `.apply`, preview, actions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pmesh_oracle as po


def test_fieldmesh_and_arraymesh_sources(cuda):
    from nbodykit_b200.lab import ArrayMesh, FieldMesh, FFTPower
    rng = np.random.RandomState(3)
    N, L = 16, 100.
    arr = rng.standard_normal((N, N, N)) + 1.0
    am = ArrayMesh(arr, BoxSize=L)
    real = am.compute(mode='real')
    np.testing.assert_allclose(real.numpy(), arr, rtol=1e-14)
    r1 = FFTPower(am, mode='1d')
    # the same through a FieldMesh wrapping the real field, and through a raw Field (fftpower.py:708-713)
    r2 = FFTPower(FieldMesh(real), mode='1d')
    r3 = FFTPower(real, mode='1d')
    o = po.power_from_complex(po.r2c(arr), None, N, L, mode='1d')
    for r in (r1, r2, r3):
        assert np.array_equal(r.power['modes'], np.squeeze(o['modes']))
        np.testing.assert_allclose(r.power['power'].real, np.squeeze(o['power']).real, rtol=1e-10)
    # complex input to ArrayMesh: irfftn(c) * N^3 (source/mesh/array.py:36-37)
    c = np.fft.rfftn(arr) / arr.size
    am2 = ArrayMesh(c, BoxSize=L)
    np.testing.assert_allclose(am2.compute(mode='real').numpy(), arr, rtol=1e-10, atol=1e-12)
    # the wrapped field is never modified
    before = real.numpy().copy()
    FieldMesh(real).compute(mode='complex')
    np.testing.assert_array_equal(real.numpy(), before)


def test_apply_user_callbacks_and_filters(cuda):
    """base/mesh.py:118-176 contract: func(x, v) with x the coordinate list for `kind`"""
    from nbodykit_b200.lab import ArrayMesh
    from nbodykit_b200.base.mesh import MeshFilter
    rng = np.random.RandomState(4)
    N, L = [8, 16, 8], [10., 40., 20.]
    arr = rng.standard_normal(N)
    mesh = ArrayMesh(arr, BoxSize=L)
    R = 3.0

    def gauss(k, v):
        k2 = sum(ki ** 2 for ki in k)
        return v * np.exp(-0.5 * k2 * R ** 2)
    out = mesh.apply(gauss, kind='wavenumber', mode='complex').compute(mode='complex').numpy()
    k = po.k_coords(N, L, 'f4')
    want = po.r2c(arr) * np.exp(-0.5 * (k[0] ** 2 + k[1] ** 2 + k[2] ** 2) * R ** 2)
    np.testing.assert_allclose(out, want, rtol=1e-6, atol=1e-12)

    class Circ(MeshFilter):
        kind = 'circular'
        mode = 'complex'

        def filter(self, w, v):
            return v * np.cos(w[0]) * np.cos(w[2])
    out = mesh.apply(Circ).compute(mode='complex').numpy()
    w = po.k_coords(N, L, 'f4', kind='circular')
    np.testing.assert_allclose(out, po.r2c(arr) * np.cos(w[0]) * np.cos(w[2]), rtol=1e-5, atol=1e-10)
    # a real-space action on index coordinates, then back to real
    out = mesh.apply(lambda i, v: v * (i[1] % 2), kind='index', mode='real').compute(mode='real').numpy()
    np.testing.assert_allclose(out, arr * (np.arange(16)[None, :, None] % 2), rtol=1e-14)


def test_compensation_action_runs_as_kernel_and_matches_host_formula(cuda):
    """mesh.compute(mode='complex') with compensated=True == r2c then the reference formula"""
    from nbodykit_b200.lab import UniformCatalog
    cat = UniformCatalog(nbar=1e-3, BoxSize=100., seed=7)
    for resampler, interlaced in [("cic", False), ("tsc", True)]:
        plain = cat.to_mesh(Nmesh=16, resampler=resampler, interlaced=interlaced, compensated=False, dtype='f8')
        comp = cat.to_mesh(Nmesh=16, resampler=resampler, interlaced=interlaced, compensated=True, dtype='f8')
        c0 = plain.compute(mode='complex').numpy()
        c1 = comp.compute(mode='complex').numpy()
        name = po.COMPENSATION[(interlaced, resampler)]
        want = po.compensate(name, po.k_coords(16, 100., 'f8', kind='circular'), c0)
        np.testing.assert_allclose(c1, want, rtol=1e-12, atol=1e-14)


def test_preview_and_attrs(cuda):
    from nbodykit_b200.lab import UniformCatalog
    cat = UniformCatalog(nbar=1e-3, BoxSize=100., seed=7)
    mesh = cat.to_mesh(Nmesh=16, dtype='f8')
    full = mesh.preview()
    real = mesh.compute(mode='real')
    np.testing.assert_allclose(full, real.numpy())
    np.testing.assert_allclose(mesh.preview(axes=(0, 1)), real.numpy().sum(axis=2))
    np.testing.assert_allclose(full.sum(), real.csum(), rtol=1e-12)             # base/tests/test_mesh.py:117-135
    for key in ['N', 'W', 'W2', 'shotnoise', 'num_per_cell', 'BoxSize', 'Nmesh', 'interlaced', 'compensated', 'resampler']:
        assert key in real.attrs
    np.testing.assert_allclose(real.cmean(), 1.0, rtol=1e-12)
    with pytest.raises(ValueError):
        mesh.compute(mode='nope')
    with pytest.raises(NotImplementedError):
        mesh.save("x")
    with pytest.raises(NotImplementedError):
        cat.to_mesh(Nmesh=16, dtype='c16')


def test_dk0_unique_bins(cuda):
    """algorithms/tests/test_fftpower.py:65-71: dk=0 -> one bin per distinct |k|, bin centres equal the mean k"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=3e-3, BoxSize=64., seed=42)
    r = FFTPower(cat, mode='1d', Nmesh=8, dk=0)
    p = r.power
    np.testing.assert_allclose(p.coords['k'], p['k'], rtol=1e-6)
    assert np.all(p['modes'] > 0)


@pytest.mark.parametrize("mode,poles", [("1d", []), ("2d", [0, 2])])
def test_fftcorr_vs_oracle(cuda, mode, poles):
    """FFTCorr (algorithms/fftcorr.py:148-176): xi = c2r(c1 c2* V, zero mode cleared) / V, binned in wrapped separation"""
    from nbodykit_b200.lab import UniformCatalog, FFTCorr
    from oracle import convpower_oracle as co
    N, L = 32, 256.
    cat = UniformCatalog(nbar=1e-3, BoxSize=L, seed=11)
    mesh = cat.to_mesh(Nmesh=N, dtype='f8', resampler='cic', compensated=True)
    r = FFTCorr(mesh, mode=mode, Nmu=4, poles=poles)
    pos, _ = po.uniform_catalog(1e-3, L, 11)
    real, _ = po.paint_field(pos, N, L, 'cic', dtype='f8')
    c = po.compensate('CompensateCICShotnoise', po.k_coords(N, L, 'f4', kind='circular'), po.r2c(real))
    p3d = c * np.conj(c)
    p3d[0, 0, 0] = 0
    p3d = p3d * L ** 3
    xi = po.c2r(p3d, N) / L ** 3
    dr = L / N
    redges = np.arange(0., 0.5 * L + dr / 2, dr)
    Nmu = 1 if mode == "1d" else 4
    res, pres = po.project_to_basis(xi, co.x_coords(N, L, 'f4'), [redges, np.linspace(0, 1, Nmu + 1)], poles=poles,
                                    hermitian_symmetric=False)
    assert np.array_equal(r.corr['modes'], np.squeeze(res[3]))
    scale = np.nanmax(np.abs(res[2]))
    np.testing.assert_allclose(np.nan_to_num(r.corr['corr'].real), np.nan_to_num(np.squeeze(res[2]).real), rtol=1e-6,
                               atol=1e-7 * scale)
    np.testing.assert_allclose(r.corr['r'], np.squeeze(res[0]), rtol=1e-6, equal_nan=True)
    if poles:
        np.testing.assert_allclose(np.nan_to_num(r.poles['corr_2'].real), np.nan_to_num(pres[1][1].real), rtol=1e-6,
                                   atol=1e-7 * scale)
    assert r.attrs['N1'] == len(pos)


def test_projected_fftpower_reference_assertions(cuda):
    """algorithms/tests/test_fftpower.py:137-155 re-run: zero mode cleared, projected power consistent with the 3-D
    power; plus a numpy restatement of the projection from the oracle's field"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower, ProjectedFFTPower
    source = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
    Nmesh = 64
    rp1 = ProjectedFFTPower(source, Nmesh=Nmesh, axes=[1])
    assert rp1.power['power'][0] == 0
    rp2 = ProjectedFFTPower(source, Nmesh=Nmesh, axes=[0, 1])
    assert rp2.power['power'][0] == 0
    rf = FFTPower(source, Nmesh=Nmesh, mode='1d')
    L = source.attrs['BoxSize'][0]
    np.testing.assert_allclose(rp1.power['power'][1:].mean() * L ** 2, rf.power['power'][1:].mean(), rtol=2 * (Nmesh / 2) ** -0.5)
    np.testing.assert_allclose(rp2.power['power'][1:].mean() * L, rf.power['power'][1:].mean(),
                               rtol=2 * (Nmesh ** 2 / 2) ** -0.5 * 10)
    # numpy restatement from the oracle's compensated field
    pos, _ = po.uniform_catalog(3e-4, 512., 42)
    real, _ = po.paint_field(pos, Nmesh, 512., 'cic', dtype='f8')
    c = po.compensate('CompensateCICShotnoise', po.k_coords(Nmesh, 512., 'f4', kind='circular'), po.r2c(real))
    r = po.c2r(c, Nmesh).sum(axis=2)                       # keep axes (0, 1)
    cp = np.fft.rfftn(r) / Nmesh ** 3
    pk = (cp * cp.conj()).real
    pk.flat[0] = 0
    kx = np.fft.fftfreq(Nmesh, 1. / (Nmesh * 2 * np.pi / 512.))
    kmag = np.sqrt(kx[:, None] ** 2 + kx[None, :Nmesh // 2 + 1] ** 2)
    W = np.full(pk.shape, 2.0); W[..., 0] = 1.0; W[..., -1] = 1.0
    dig = np.digitize(kmag.flat, rp2.edges)
    nb = len(rp2.edges) + 1
    with np.errstate(invalid='ignore', divide='ignore'):
        want = (np.bincount(dig, weights=(W * pk).flat, minlength=nb) / np.bincount(dig, weights=W.flat, minlength=nb))[1:-1] * 512. ** 2
    np.testing.assert_allclose(rp2.power['power'].real, want, rtol=1e-6, atol=1e-9 * np.nanmax(want))
    assert np.array_equal(rp2.power['modes'], np.bincount(dig, weights=W.flat, minlength=nb)[1:-1])


def test_fftcorr_unique_and_padding(cuda):
    """algorithms/tests/test_fftcorr.py: dr=0 -> one bin per distinct separation (bin centres == mean r), and a
    catalogue painted on a larger box keeps N1 / N2"""
    from nbodykit_b200.lab import UniformCatalog, FFTCorr
    source = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
    p = FFTCorr(source, mode='1d', Nmesh=32, dr=0).corr
    np.testing.assert_allclose(p.coords['r'], p['r'], rtol=1e-6)
    r = FFTCorr(source, mode='1d', BoxSize=1024, Nmesh=32)
    assert r.attrs['N1'] != 0 and r.attrs['N2'] != 0
