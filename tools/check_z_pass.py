"""r2c of long z rows (Nz = 256 .. 4096: every last-stage radix of the register-I/O z pass) against numpy.fft"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from nbodykit_b200.pmesh.pm import ParticleMesh, RealField
from nbodykit_b200.comm import SelfComm
rng = np.random.RandomState(1)
for N in ([8, 8, 256], [4, 8, 512], [4, 4, 1024], [2, 4, 2048], [4, 4, 4096]):
    for dt, tol in (("f8", 1e-13), ("f4", 2e-6)):
        real = rng.standard_normal(N).astype(dt)
        pm = ParticleMesh(BoxSize=1.0, Nmesh=N, dtype=dt, comm=SelfComm())
        f = RealField(pm); f[...] = real
        got = f.r2c().numpy()
        want = np.fft.rfftn(real.astype("f8")) / real.size
        err = np.abs(got - want).max() / (np.sqrt((np.abs(want) ** 2).mean()) * np.log2(real.size))
        print(N, dt, "rel err %.2e" % err, "OK" if err <= tol else "FAIL")
