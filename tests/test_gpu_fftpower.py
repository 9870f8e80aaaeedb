"""
End-to-end parity: the nbodykit_b200 FFTPower API on the GPU against the CPU oracle's restatement of
the whole reference flow (paint -> r2c -> compensate -> |delta|^2 V -> project_to_basis).
Mode counts / N1 / N2 bit-exact; k, mu, P within 1e-5 relative (BASELINE north_star tolerance).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pmesh_oracle as po

RTOL = 1e-5


def _compare(r, o, mode):
    assert np.array_equal(r.power['modes'], np.squeeze(o['modes']))
    np.testing.assert_allclose(r.power['k'], np.squeeze(o['k']), rtol=RTOL, equal_nan=True)
    P, Po = r.power['power'], np.squeeze(o['power'])
    scale = np.nanmax(np.abs(Po))
    assert np.array_equal(np.isnan(P.real), np.isnan(Po.real))
    np.testing.assert_allclose(np.nan_to_num(P.real), np.nan_to_num(Po.real), rtol=RTOL, atol=RTOL * 1e-3 * scale)
    np.testing.assert_allclose(np.nan_to_num(P.imag), np.nan_to_num(Po.imag), rtol=RTOL, atol=RTOL * 1e-3 * scale)
    if mode == '2d':
        np.testing.assert_allclose(r.power['mu'], o['mu'], rtol=RTOL, atol=1e-7, equal_nan=True)


def test_config1_uniform_cic_1d(cuda):
    """BASELINE config 1: UniformCatalog(nbar=1e5/1024^3, BoxSize=1024, seed=42) -> 64^3 CIC, mode='1d'"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=1e5 / 1024. ** 3, BoxSize=1024., seed=42)
    assert cat.csize == 99886
    r = FFTPower(cat, mode='1d', Nmesh=64)
    pos, _ = po.uniform_catalog(1e5 / 1024. ** 3, 1024., 42)
    o = po.fftpower(pos, 64, 1024., mode='1d', resampler='cic', compensated=True, dtype='f8')
    assert r.attrs['N1'] == 99886 and r.attrs['N2'] == 99886
    np.testing.assert_allclose(r.attrs['shotnoise'], 1024. ** 3 / 99886, rtol=1e-12)
    assert r.power.shape == (32,)
    _compare(r, o, '1d')


def test_doc_known_answer_N1_96(cuda):
    """docs/source/results/algorithms/fftpower.ipynb:313-315"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=100, BoxSize=1.0, seed=42)
    r = FFTPower(cat, mode='1d', Nmesh=32)
    assert r.attrs['N1'] == 96
    np.testing.assert_allclose(r.attrs['shotnoise'], 1. / 96, rtol=1e-12)


@pytest.mark.parametrize("resampler,interlaced,dtype", [("tsc", True, "f4"), ("tsc", False, "f8"), ("cic", True, "f8"),
                                                        ("pcs", False, "f4")])
def test_mesh_variants_2d_poles(cuda, resampler, interlaced, dtype):
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
    mesh = cat.to_mesh(Nmesh=32, resampler=resampler, interlaced=interlaced, compensated=True, dtype=dtype)
    r = FFTPower(mesh, mode='2d', Nmu=5, poles=[0, 2, 4], los=[0, 0, 1])
    pos, _ = po.uniform_catalog(3e-4, 512., 42)
    o = po.fftpower(pos, 32, 512., mode='2d', resampler=resampler, interlaced=interlaced, compensated=True,
                    dtype=dtype, Nmu=5, poles=[0, 2, 4])
    tol = RTOL
    assert np.array_equal(r.power['modes'], o['modes'])
    assert np.array_equal(r.poles['modes'], o['poles_modes'])
    np.testing.assert_allclose(r.power['k'], o['k'], rtol=tol, equal_nan=True)
    # (k,mu) wedges and ell > 0 multipoles are sums with cancellations: their error is measured against the
    # amplitude of the monopole (f4 meshes carry 6e-8 rounding per mode through both pipelines)
    scale = np.nanmax(np.abs(o['poles_power'][0]))
    atol = (1e-5 if dtype == "f4" else 1e-7) * scale
    np.testing.assert_allclose(np.nan_to_num(r.power['power'].real), np.nan_to_num(o['power'].real), rtol=tol, atol=atol)
    np.testing.assert_allclose(r.poles['power_0'].real, o['poles_power'][0].real, rtol=tol, equal_nan=True)
    for i, ell in enumerate([0, 2, 4]):
        np.testing.assert_allclose(np.nan_to_num(r.poles['power_%d' % ell].real), np.nan_to_num(o['poles_power'][i].real),
                                   rtol=tol, atol=atol)
    # reference identity (test_fftpower.py:49-61): monopole from P(k,mu) equals poles['power_0']
    modes_1d = r.power['modes'].sum(axis=-1)
    mono = np.nansum(r.power['power'].real * r.power['modes'], axis=-1) / modes_1d
    assert np.array_equal(modes_1d, r.poles['modes'])
    np.testing.assert_allclose(mono, r.poles['power_0'].real, rtol=1e-10, equal_nan=True)


def test_weighted_and_selection(cuda):
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
    rng = np.random.RandomState(9)
    w = rng.uniform(0.5, 2.0, size=cat.size)
    sel = rng.uniform(size=cat.size) < 0.7
    cat['Weight'] = w
    cat['Selection'] = sel
    r = FFTPower(cat, mode='1d', Nmesh=32)
    pos, _ = po.uniform_catalog(3e-4, 512., 42)
    o = po.fftpower(pos[sel], 32, 512., mode='1d', weight=w[sel], dtype='f8')
    assert r.attrs['N1'] == int(sel.sum())
    np.testing.assert_allclose(r.attrs['shotnoise'], o['attrs']['shotnoise'], rtol=1e-12)
    _compare(r, o, '1d')


def test_cross_power_and_device_resident_columns(cuda):
    """second != first; columns handed over as CUDA tensors (no host round trip)"""
    import torch
    from nbodykit_b200.lab import ArrayCatalog, FFTPower
    rng = np.random.RandomState(1)
    p1 = rng.uniform(0, 256., size=(20000, 3)).astype('f4')
    p2 = rng.uniform(0, 256., size=(15000, 3)).astype('f4')
    c1 = ArrayCatalog({'Position': torch.from_numpy(p1).cuda()}, BoxSize=256.)
    c2 = ArrayCatalog({'Position': torch.from_numpy(p2).cuda()}, BoxSize=256.)
    r = FFTPower(c1, mode='1d', Nmesh=32, second=c2)
    assert r.attrs['shotnoise'] == 0 and r.attrs['N1'] == 20000 and r.attrs['N2'] == 15000
    # oracle: two separate fields
    N, L = 32, 256.
    f1, _ = po.paint_field(p1, N, L, 'cic', dtype='f8')
    f2, _ = po.paint_field(p2, N, L, 'cic', dtype='f8')
    wc = po.k_coords(N, L, 'f4', kind='circular')
    k1 = po.compensate('CompensateCICShotnoise', wc, po.r2c(f1))
    k2 = po.compensate('CompensateCICShotnoise', wc, po.r2c(f2))
    o = po.power_from_complex(k1, k2, N, L, mode='1d')
    _compare(r, o, '1d')


def test_empty_selection_gives_unit_field(cuda):
    """source/mesh/tests/test_catalogmesh.py:27-43"""
    import warnings
    from nbodykit_b200.lab import UniformCatalog
    cat = UniformCatalog(nbar=3e-4, BoxSize=64., seed=42)
    cat['Selection'] = np.zeros(cat.size, dtype=bool)
    mesh = cat.to_mesh(Nmesh=8, dtype='f8')
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        real = mesh.compute(mode='real')
    np.testing.assert_allclose(real.numpy(), 1.0)
    assert real.attrs['N'] == 0


def test_save_load_roundtrip(cuda, tmp_path):
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    cat = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
    r = FFTPower(cat, mode='2d', Nmesh=32, Nmu=3, poles=[0, 2])
    fn = str(tmp_path / "r.json")
    r.save(fn)
    r2 = FFTPower.load(fn)
    for name in r.power.variables:
        np.testing.assert_array_equal(r.power[name], r2.power[name])
    for name in r.poles.variables:
        np.testing.assert_array_equal(r.poles[name], r2.poles[name])
    assert r2.attrs['N1'] == r.attrs['N1']


def test_reference_assertions_chi2(cuda):
    """algorithms/tests/test_fftpower.py:12-44: compensated CIC/TSC shot-noise spectra have reduced chi^2 < 1"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    for resampler in ['cic', 'tsc']:
        source = UniformCatalog(nbar=3e-4, BoxSize=512., seed=42)
        mesh = source.to_mesh(resampler=resampler, Nmesh=64, compensated=True)
        r = FFTPower(mesh, mode='1d', kmin=0.02)
        Pk = r.power['power'].real
        err = (2 * Pk ** 2 / r.power['modes']) ** 0.5
        residual = (Pk - r.attrs['shotnoise']) / err
        red_chi2 = (residual ** 2).sum() / len(Pk)        # "should be about 0.5-0.6"
        assert red_chi2 < 1.0, (resampler, red_chi2)


def test_reference_assertion_tsc_interlacing(cuda):
    """source/mesh/tests/test_catalogmesh.py:12-23 -- pins the interlacing sign / shift convention"""
    from nbodykit_b200.lab import UniformCatalog, FFTPower
    source = UniformCatalog(nbar=3e-2, BoxSize=512., seed=42)
    mesh = source.to_mesh(resampler='tsc', Nmesh=64, interlaced=True, compensated=True)
    r = FFTPower(mesh, mode='1d', kmin=0.02)
    np.testing.assert_allclose(r.power['power'][5:].real, 1. / 3e-2, rtol=1e-1)


def test_host_resident_streamed_paint_equals_device_resident(cuda):
    """host columns are staged in paint_chunk_size pieces overlapped with the scatter (source/mesh/tests/
    test_catalogmesh.py:47-60: the result must not depend on the chunk size)"""
    import torch
    from nbodykit_b200 import set_options
    from nbodykit_b200.lab import ArrayCatalog
    rng = np.random.RandomState(2)
    pos = rng.uniform(0, 100., size=(500000, 3)).astype('f4')
    dev = ArrayCatalog({'Position': torch.from_numpy(pos).cuda()}, BoxSize=100.)
    want = dev.to_mesh(Nmesh=32, resampler='tsc', dtype='f8').compute(mode='real').numpy()
    for chunk in (50000, 33333):
        with set_options(paint_chunk_size=chunk):
            host = ArrayCatalog({'Position': pos}, BoxSize=100.)
            got = host.to_mesh(Nmesh=32, resampler='tsc', dtype='f8').compute(mode='real')
            assert got.attrs['N'] == len(pos)
            np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=1e-7 * want.max())
            pinned = torch.from_numpy(pos).pin_memory()
            got2 = ArrayCatalog({'Position': pinned}, BoxSize=100.).to_mesh(Nmesh=32, resampler='tsc', interlaced=True,
                                                                            dtype='f8').compute(mode='real')
    want_i = dev.to_mesh(Nmesh=32, resampler='tsc', interlaced=True, dtype='f8').compute(mode='real').numpy()
    np.testing.assert_allclose(got2.numpy(), want_i, rtol=0, atol=1e-7 * want_i.max())
