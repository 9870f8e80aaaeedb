import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pmesh_oracle as po
from nbodykit_b200.comm import SelfComm
from nbodykit_b200.lab import ArrayCatalog, FFTPower
N, L = 256, 512.
rng = np.random.RandomState(12)
pos = rng.uniform(size=(2000000, 3)) * L
r = FFTPower(ArrayCatalog({'Position': pos}, comm=SelfComm(), BoxSize=L), mode='1d', Nmesh=N, dk=0)
kedges = r.power.edges['k']
print("edges dtype", kedges.dtype, len(kedges))
k3 = po.k_coords(N, L, 'f4')
ones = np.ones((N, N, N // 2 + 1), dtype='c16')
for name, e in (("as-is", kedges), ("f8", kedges.astype('f8'))):
    res, _ = po.project_to_basis(ones, k3, [e, np.linspace(-1, 1, 2)])
    m = np.squeeze(res[3])
    d = np.nonzero(m != r.power['modes'])[0]
    print(name, "mismatching bins:", len(d), d[:10], m[d[:10]], r.power['modes'][d[:10]], kedges[d[:10]], kedges[d[:10] + 1])
print("sum modes gpu", r.power['modes'].sum())
