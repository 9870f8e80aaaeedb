"""scratch: locate the FFTCorr mismatch (field level vs binning level)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pmesh_oracle as po, convpower_oracle as co
from nbodykit_b200.lab import UniformCatalog, FFTCorr
from nbodykit_b200.algorithms.fftpower import project_to_basis_device
from nbodykit_b200.pmesh.pm import ComplexField


N, L = 32, 256.
cat = UniformCatalog(nbar=1e-3, BoxSize=L, seed=11)
mesh = cat.to_mesh(Nmesh=N, dtype='f8', resampler='cic', compensated=True)
r = FFTCorr(mesh, mode='1d', Nmu=4)
pos, _ = po.uniform_catalog(1e-3, L, 11)
real, _ = po.paint_field(pos, N, L, 'cic', dtype='f8')
c = po.compensate('CompensateCICShotnoise', po.k_coords(N, L, 'f4', kind='circular'), po.r2c(real))
p3d = c * np.conj(c); p3d[0, 0, 0] = 0; p3d = p3d * L ** 3
xi = po.c2r(p3d, N) / L ** 3
# field-level: product's complex field
cf = mesh.compute(mode='complex')
got_c = cf.numpy()
print("complex field max rel diff", np.abs(got_c - c).max() / np.abs(c).max())
p = cf.copy() if hasattr(cf, 'copy') else cf
dr = L / N
redges = np.arange(0., 0.5 * L + dr / 2, dr)
res, _ = po.project_to_basis(xi, co.x_coords(N, L, 'f4'), [redges, np.linspace(0, 1, 2)], poles=[], hermitian_symmetric=False)
print("modes equal", np.array_equal(r.corr['modes'], np.squeeze(res[3])))
print("r   ", r.corr['r'][:6], np.squeeze(res[0])[:6])
print("corr", r.corr['corr'][:6], np.squeeze(res[2])[:6].real)
print("ratio", (r.corr['corr'] / np.squeeze(res[2]).real)[:8])
