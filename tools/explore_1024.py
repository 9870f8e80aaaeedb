"""Scratch: where does the 1e9-particle / 1024^3 single-GPU FFTPower stand?  (stage times, memory, both orders)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodykit_b200 import _lib
from nbodykit_b200.comm import SelfComm
from nbodykit_b200.cosmology import NoWiggleEHPower
from nbodykit_b200.lab import ArrayCatalog, FFTPower
from nbodykit_b200.source.catalog.lognormal import LogNormalCatalog


def run(pos, Nmesh, Box, label, reps=3, **kw):
    cat = ArrayCatalog({'Position': pos}, comm=SelfComm(), BoxSize=Box)
    for i in range(reps + 1):
        _lib.profiler.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = FFTPower(cat, mode='1d', Nmesh=Nmesh, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = _lib.profiler.stop()
        if i:
            print("%s: %.2f ms  " % (label, dt * 1e3) + " ".join("%s=%.2f" % (k, sum(v)) for k, v in sorted(st.items())),
                  flush=True)
    print("   max mem %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
    return r


def main():
    npart = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
    Nmesh = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    Box = 2048.0
    gen = Nmesh // 2
    torch.cuda.set_device(0)
    t0 = time.time()
    cat = LogNormalCatalog(Plin=NoWiggleEHPower(), nbar=npart / Box ** 3, BoxSize=Box, Nmesh=gen, bias=2.0, seed=42,
                           comm=SelfComm())
    pos = cat['Position'].compute()
    del cat
    torch.cuda.synchronize()
    print("generated %d particles in %.1f s, max mem %.1f GB" % (pos.shape[0], time.time() - t0,
                                                                 torch.cuda.max_memory_allocated() / 1e9), flush=True)
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    r1 = run(pos, Nmesh, Box, "sorted  ")
    g = torch.Generator(device=pos.device)
    g.manual_seed(45)
    perm = torch.randperm(pos.shape[0], device=pos.device, generator=g)
    pp = pos[perm].contiguous()
    del perm
    r2 = run(pp, Nmesh, Box, "permuted")
    print("modes equal:", np.array_equal(r1.power['modes'], r2.power['modes']),
          " max rel dP:", float(np.nanmax(np.abs(r1.power['power'].real / r2.power['power'].real - 1))))
    del pp
    os.environ["NBK_PAINT_BUCKET"] = "global"


if __name__ == "__main__":
    main()
