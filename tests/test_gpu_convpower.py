"""ConvolvedFFTPower / FKP / multi-species paint on the GPU against the NumPy restatement (oracle/convpower_oracle.py)
and against the reference's own identities (algorithms/tests/test_conv_power.py, convpower/tests/test_catalogmesh.py,
source/mesh/tests/test_species.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pmesh_oracle as po, convpower_oracle as co

NBAR = 3e-4


def _fkp(seed_d=42, seed_r=84, P0=1e4, box=512., shift=(1000., -300., 700.)):
    from nbodykit_b200.lab import UniformCatalog, FKPCatalog
    d = UniformCatalog(nbar=NBAR, BoxSize=box, seed=seed_d)
    r = UniformCatalog(nbar=10 * NBAR, BoxSize=box, seed=seed_r)
    for c in (d, r):
        c['Position'] = c['Position'] + np.array(shift)      # an off-origin survey volume
        c['NZ'] = NBAR
    rng = np.random.RandomState(5)
    d['Weight'] = rng.uniform(0.8, 1.2, size=d.size)
    return FKPCatalog(d, r, P0=P0), d, r


def test_ylm_kernels_match_reference_values(cuda):
    """nbk_ylm_mul_real / nbk_ylm_mul_complex_acc reproduce the reference's get_real_Ylm (golden values)"""
    import ctypes
    import os
    import torch
    from nbodykit_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ylm_reference.npz"))
    v = g["vec"]
    for l in range(5):
        for m in range(-l, l + 1):
            np.testing.assert_allclose(co.real_ylm(l, m, v[:, 0], v[:, 1], v[:, 2]) * np.ones(64), g["Y_%d_%d" % (l, m)],
                                       rtol=1e-12, atol=1e-14)
            assert abs(float(co.real_ylm(l, m, 0., 0., 0.)) - float(g["Y0_%d_%d" % (l, m)])) < 1e-14
    # the kernel: a 4^3 real field of ones times Y_lm(xhat) with a large offset -> compare to the oracle grid
    N, L = [4, 8, 4], [8., 8., 4.]
    off = np.array([100., -50., 25.])
    xs = co.x_coords(N, L, "f8")
    xg = [x + off[i] for i, x in enumerate(xs)]
    rn = np.sqrt(sum(x ** 2 for x in xg))
    ones = torch.ones(tuple(N), dtype=torch.float64, device="cuda")
    out = torch.empty_like(ones)
    Lb = _lib.lib()
    for (l, m) in [(0, 0), (1, -1), (2, 0), (2, 2), (3, -2), (4, 0), (4, 3), (6, -5), (8, 8)]:
        _lib.check(Lb.nbk_ylm_mul_real(ctypes.c_void_p(ones.data_ptr()), ctypes.c_void_p(out.data_ptr()), 8, l, m,
                                       _lib.iarr(N), _lib.darr(L), _lib.darr(off), 0, N[0], None))
        torch.cuda.synchronize()
        want = co.real_ylm(l, m, xg[0] / rn, xg[1] / rn, xg[2] / rn) * np.ones(N)
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-11, atol=1e-13)
    # Fourier side, including khat := 0 at k = 0
    ks = po.k_coords(N, L, "f8")
    kn = np.sqrt(sum(k ** 2 for k in ks))
    kn[kn == 0] = np.inf
    shape = (N[0], N[1], N[2] // 2 + 1)
    c = torch.ones(shape, dtype=torch.complex128, device="cuda") * (1 + 2j)
    for (l, m) in [(2, 1), (4, 0), (4, -4)]:
        acc = torch.zeros(shape, dtype=torch.complex128, device="cuda")
        _lib.check(Lb.nbk_ylm_mul_complex_acc(ctypes.c_void_p(acc.data_ptr()), ctypes.c_void_p(c.data_ptr()), 8, l, m,
                                              _lib.iarr(N), _lib.darr(L), 0, 0, N[0], None))
        torch.cuda.synchronize()
        want = (1 + 2j) * co.real_ylm(l, m, ks[0] / kn, ks[1] / kn, ks[2] / kn) * np.ones(shape)
        np.testing.assert_allclose(acc.cpu().numpy(), want, rtol=1e-11, atol=1e-13)


def test_multiple_species_paint_is_sum_of_single_paints(cuda):
    """source/mesh/tests/test_species.py:50-75"""
    from nbodykit_b200.lab import UniformCatalog, MultipleSpeciesCatalog
    s1 = UniformCatalog(nbar=3e-4, BoxSize=256., seed=42)
    s2 = UniformCatalog(nbar=1e-4, BoxSize=256., seed=84)
    cat = MultipleSpeciesCatalog(['data', 'randoms'], s1, s2)
    mesh = cat.to_mesh(Nmesh=32, BoxSize=256., dtype='f8')
    combined = mesh.compute(mode='real')
    a = s1.to_mesh(Nmesh=32, BoxSize=256., dtype='f8').to_real_field(normalize=False)
    b = s2.to_mesh(Nmesh=32, BoxSize=256., dtype='f8').to_real_field(normalize=False)
    norm = a.attrs['num_per_cell'] + b.attrs['num_per_cell']
    np.testing.assert_allclose(combined.numpy(), (a.numpy() + b.numpy()) / norm, atol=1e-5)
    assert combined.attrs['N'] == s1.csize + s2.csize
    W1, W2 = combined.attrs['data.W'], combined.attrs['randoms.W']
    want = (W1 / (W1 + W2)) ** 2 * combined.attrs['data.shotnoise'] + (W2 / (W1 + W2)) ** 2 * combined.attrs['randoms.shotnoise']
    np.testing.assert_allclose(combined.attrs['shotnoise'], want)


def test_fkp_field_is_data_minus_alpha_randoms(cuda):
    """convpower/tests/test_catalogmesh.py:15-83"""
    fkp, d, r = _fkp()
    mesh = fkp.to_mesh(Nmesh=32, dtype='f8')
    real = mesh.compute(mode='real')
    alpha = real.attrs['alpha']
    np.testing.assert_allclose(alpha, float(d['Weight'].sum()) / r.csize)
    C, L = mesh.attrs['BoxCenter'], mesh.attrs['BoxSize']
    wfd = 1. / (1 + 1e4 * NBAR)
    F, alpha_o = co.fkp_field(np.asarray(d['Position']), np.asarray(r['Position']),
                              (np.asarray(d['Weight']), wfd * np.ones(d.size)), (np.ones(r.size), wfd * np.ones(r.size)),
                              32, L, C)
    np.testing.assert_allclose(real.numpy(), F, rtol=0, atol=1e-7 * np.abs(F).max())
    assert real.attrs['data.N'] == d.csize and real.attrs['randoms.N'] == r.csize


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
def test_convolved_power_vs_oracle(cuda, resampler):
    from nbodykit_b200.lab import ConvolvedFFTPower
    fkp, d, r = _fkp()
    mesh = fkp.to_mesh(Nmesh=32, dtype='f8', resampler=resampler)
    res = ConvolvedFFTPower(mesh, poles=[0, 2, 4], dk=0.02)
    C, L = mesh.attrs['BoxCenter'], mesh.attrs['BoxSize']
    wfd = 1. / (1 + 1e4 * NBAR)
    o = co.convpower(np.asarray(d['Position']), np.asarray(r['Position']),
                     (np.asarray(d['Weight']), wfd * np.ones(d.size)), (np.ones(r.size), wfd * np.ones(r.size)),
                     NBAR * np.ones(d.size), NBAR * np.ones(r.size), 32, L, C, [0, 2, 4], resampler=resampler, dk=0.02)
    np.testing.assert_allclose(res.attrs['alpha'], o['alpha'], rtol=1e-12)
    np.testing.assert_allclose(res.attrs['data.norm'], o['data_norm'], rtol=1e-12)
    np.testing.assert_allclose(res.attrs['randoms.norm'], o['randoms_norm'], rtol=1e-12)
    np.testing.assert_allclose(res.attrs['shotnoise'], o['shotnoise'], rtol=1e-12)
    assert np.array_equal(res.poles['modes'], o['modes'])
    np.testing.assert_allclose(res.poles['k'], o['k'], rtol=1e-6, equal_nan=True)
    scale = np.nanmax(np.abs(o['power_0']))
    for ell in (0, 2, 4):
        got, want = res.poles['power_%d' % ell], o['power_%d' % ell]
        assert got.dtype == np.complex64                       # fkp.py:451
        np.testing.assert_allclose(np.nan_to_num(got.real), np.nan_to_num(want.real), rtol=1e-5, atol=2e-6 * scale)
    # reference identities (test_conv_power.py:173-180): norms and the shot-noise formula
    S_d = float((np.asarray(d['Weight']) ** 2).sum()) * wfd ** 2
    S_r = r.csize * wfd ** 2
    np.testing.assert_allclose(res.attrs['shotnoise'], (S_d + res.attrs['alpha'] ** 2 * S_r) / res.attrs['randoms.norm'])
    # wedges from multipoles + JSON round trip
    pkmu = res.to_pkmu(np.linspace(0, 1, 4), 4)
    assert pkmu.shape == (len(res.poles['k']), 3)


def test_convolved_power_complex_mesh_odd_multipoles(cuda):
    """dtype='c16' (the reference default): every multipole equals the sum over ALL modes of a full complex mesh
    (oracle: fftn + project_to_basis(hermitian_symmetric=False)) in every bin; with an 'f8' mesh the odd multipoles
    keep the reference's Hermitian fold (documented there as incorrect)"""
    from nbodykit_b200.lab import ConvolvedFFTPower
    fkp, d, r = _fkp()
    mesh = fkp.to_mesh(Nmesh=32)                       # default dtype, as in the reference
    assert mesh.complex_mesh
    res = ConvolvedFFTPower(mesh, poles=[0, 1, 2, 3], dk=0.02)
    C, L = mesh.attrs['BoxCenter'], mesh.attrs['BoxSize']
    wfd = 1. / (1 + 1e4 * NBAR)
    args = (np.asarray(d['Position']), np.asarray(r['Position']),
            (np.asarray(d['Weight']), wfd * np.ones(d.size)), (np.ones(r.size), wfd * np.ones(r.size)),
            NBAR * np.ones(d.size), NBAR * np.ones(r.size), 32, L, C, [0, 1, 2, 3])
    o = co.convpower_full(*args, dk=0.02)
    assert np.array_equal(res.poles['modes'], o['modes'])
    np.testing.assert_allclose(res.poles['k'], o['k'], rtol=1e-6, equal_nan=True)
    scale = np.nanmax(np.abs(o['power_0']))
    # EVERY bin, including the Nyquist shell: modes with j_x or j_y = -N/2 have mirror partners that carry the SAME
    # label -N/2 on a full mesh, so Y_lm(khat) of the pair is not related by parity; the mirror accumulator
    # (nbk_ylm_mul_complex_acc2 -> nbk_power_bin2) reproduces exactly that
    for ell in (0, 1, 2, 3):
        got, want = res.poles['power_%d' % ell], o['power_%d' % ell]
        np.testing.assert_allclose(np.nan_to_num(got.real), np.nan_to_num(want.real), rtol=1e-5, atol=2e-6 * scale)
        np.testing.assert_allclose(np.nan_to_num(got.imag), np.nan_to_num(want.imag), rtol=1e-5, atol=2e-6 * scale)
    # the odd multipoles of a survey-like geometry are imaginary and do not vanish
    assert np.nanmax(np.abs(res.poles['power_1'].imag)) > 1e-3 * scale
    assert np.nanmax(np.abs(res.poles['power_1'].real[:-1])) < 1e-5 * scale    # (the Nyquist-shell bin is not purely imaginary)
    # Hermitian mesh: the reference's own (Hermitian) fold -> the restatement with hermitian_symmetric=True
    res8 = ConvolvedFFTPower(fkp.to_mesh(Nmesh=32, dtype='f8'), poles=[1], dk=0.02)
    o8 = co.convpower(*args[:-1], [1], dk=0.02)
    for part in ('real', 'imag'):
        np.testing.assert_allclose(np.nan_to_num(getattr(res8.poles['power_1'], part)),
                                   np.nan_to_num(getattr(o8['power_1'], part)), rtol=1e-5, atol=2e-6 * scale)


def test_convolved_power_errors_and_io(cuda, tmp_path):
    from nbodykit_b200.lab import ConvolvedFFTPower
    fkp, d, r = _fkp()
    with pytest.raises(TypeError):
        ConvolvedFFTPower(d, poles=[0])
    with pytest.raises(ValueError):
        ConvolvedFFTPower(fkp, poles=[0], Nmesh=32, use_fkp_weights=True)
    res = ConvolvedFFTPower(fkp, poles=0, Nmesh=32)
    assert res.attrs['poles'] == [0] and 'power_0' in res.poles.variables
    fn = str(tmp_path / "conv.json")
    res.save(fn)
    back = ConvolvedFFTPower.load(fn)
    np.testing.assert_array_equal(back.poles['power_0'], res.poles['power_0'])
    # data.norm and randoms.norm must agree within 5 %
    bad, _, _ = _fkp()
    bad['randoms/NZ'] = 2 * NBAR
    with pytest.raises(ValueError):
        ConvolvedFFTPower(bad, poles=[0], Nmesh=32)
