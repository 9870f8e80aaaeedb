"""
N > 1 host logic on CPU: two processes over the gloo backend (127.0.0.1).  Covers the communicator shim,
rank-count invariance of the catalogue RNG, and particle routing (pm.decompose / Layout.exchange with ghost
duplication), checked against the oracle's slab painting.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbodykit_b200.comm import TorchComm
        ret[rank] = fn(TorchComm(), rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _collectives(comm, rank, world):
    out = {}
    out["allreduce"] = comm.allreduce(rank + 1)
    out["allreduce_f"] = comm.allreduce(0.5 * (rank + 1))
    out["allreduce_arr"] = comm.allreduce(np.arange(3.) * (rank + 1)).tolist()
    out["allgather"] = comm.allgather({"r": rank})
    out["bcast"] = comm.bcast("hello" if rank == 0 else None)
    out["alltoall"] = comm.alltoall([10 * rank + d for d in range(world)])
    t = torch.arange(4, dtype=torch.float64) + rank
    comm.allreduce_tensor(t)
    out["tensor"] = t.tolist()
    send = torch.arange(6, dtype=torch.float32).reshape(6, 1) + 100 * rank
    splits_in = [2, 4] if rank == 0 else [5, 1]
    splits_out = [2, 5] if rank == 0 else [4, 1]
    recv = torch.empty((sum(splits_out), 1))
    comm.all_to_all_single(recv, send, splits_out, splits_in)
    out["a2a"] = recv.reshape(-1).tolist()
    return out


def test_comm_shim_collectives():
    r0, r1 = _run(_collectives)
    assert r0["allreduce"] == 3 and r1["allreduce_f"] == 1.5
    assert r0["allreduce_arr"] == [0., 3., 6.]
    assert r0["allgather"] == [{"r": 0}, {"r": 1}] and r1["bcast"] == "hello"
    assert r0["alltoall"] == [0, 10] and r1["alltoall"] == [1, 11]
    assert r1["tensor"] == [1., 3., 5., 7.]
    assert r0["a2a"] == [0., 1., 100., 101., 102., 103., 104.]
    assert r1["a2a"] == [2., 3., 4., 5., 105.]


def _uniform(comm, rank, world):
    from nbodykit_b200.lab import UniformCatalog
    cat = UniformCatalog(nbar=2.5e5, BoxSize=1.0, seed=42, comm=comm)
    return cat['Position'].compute(), cat['Velocity'].compute(), cat.csize


def test_uniform_catalog_rank_count_invariance():
    """source/catalog/tests/test_uniform.py:8-21: gathered 2-rank result == 1-rank result, bit for bit"""
    from oracle import pmesh_oracle as po
    parts = _run(_uniform)
    pos = np.concatenate([p[0] for p in parts])
    vel = np.concatenate([p[1] for p in parts])
    want_p, want_v = po.uniform_catalog(2.5e5, 1.0, 42)
    assert parts[0][2] == len(want_p)
    np.testing.assert_array_equal(pos, want_p)
    np.testing.assert_array_equal(vel, want_v)


def _route(comm, rank, world):
    from nbodykit_b200.pmesh.pm import ParticleMesh
    N, L = [16, 8, 8], [32., 8., 8.]
    pm = ParticleMesh(BoxSize=L, Nmesh=N, dtype='f8', comm=comm)
    rng = np.random.RandomState(100 + rank)
    pos = rng.uniform(-5, 40, size=(3000, 3))          # also outside the box: periodic routing
    mass = rng.uniform(size=3000)
    out = {}
    for name, smoothing in [("cic", 1.0), ("tsc", 1.5), ("tsc_interlaced", 3.0)]:
        lay = pm.decompose(torch.from_numpy(pos), smoothing=smoothing)
        p = lay.exchange(torch.from_numpy(pos)).numpy()
        m = lay.exchange(torch.from_numpy(mass)).numpy()
        out[name] = (p, m, lay.recvlength, pm.x_start, pm.x_n)
    return pos, mass, out


def test_decompose_exchange_routes_ghosts():
    """after routing, painting each rank's received particles into ITS slab (dropping out-of-slab stencil
    points) reassembles exactly the single-rank mesh -- pmesh decompose/exchange semantics"""
    from oracle import pmesh_oracle as po
    N, L = [16, 8, 8], [32., 8., 8.]
    res = _run(_route)
    pos = np.concatenate([r[0] for r in res])
    mass = np.concatenate([r[1] for r in res])
    for name, resampler, shift in [("cic", "cic", 0.0), ("tsc", "tsc", 0.0), ("tsc_interlaced", "tsc", 0.5)]:
        full = po.paint(pos, mass, N, L, resampler, shift)
        slabs = []
        nrecv = 0
        for r in res:
            p, m, n, x0, xn = r[2][name]
            assert len(p) == n
            nrecv += n
            slabs.append(po.paint(p, m, N, L, resampler, shift, x_start=x0, x_n=xn))
        np.testing.assert_allclose(np.concatenate(slabs, axis=0), full, rtol=0, atol=1e-12 * full.max())
        assert len(pos) <= nrecv < 1.8 * len(pos)      # ghosts are duplicated, nothing is lost


def test_particle_mesh_slab_layout():
    from nbodykit_b200.comm import SelfComm
    from nbodykit_b200.pmesh.pm import ParticleMesh

    class Fake(SelfComm):
        def __init__(self, rank, size):
            self.rank, self.size = rank, size
    pm = ParticleMesh(BoxSize=1., Nmesh=[32, 16, 8], dtype='f4', comm=Fake(3, 4))
    assert (pm.x_start, pm.x_n, pm.y_start, pm.y_n) == (24, 8, 12, 4)
    assert pm.real_shape == (8, 16, 8) and pm.complex_shape == (4, 32, 5) and pm.transposed
    k = pm.create_coords("complex")
    assert k[0].shape == (1, 32, 1) and k[1].shape == (4, 1, 1) and k[2].shape == (1, 1, 5)
    assert k[0].dtype == np.float32
    np.testing.assert_allclose(k[1].ravel(), 2 * np.pi * np.array([-4, -3, -2, -1]), rtol=1e-6)
    with pytest.raises(ValueError):
        ParticleMesh(BoxSize=1., Nmesh=[30, 16, 8], dtype='f4', comm=Fake(0, 4))
    # complex-dtype meshes keep all N^3 modes (single GPU; several GPUs -> dtype='f8')
    cm = ParticleMesh(BoxSize=1., Nmesh=8, dtype='c16', comm=Fake(0, 1))
    assert cm.cplx and cm.complex_shape == (8, 8, 8) and cm.real_shape == (8, 8, 8) and cm.typestr == 'f8'
    with pytest.raises(NotImplementedError):
        ParticleMesh(BoxSize=1., Nmesh=8, dtype='c16', comm=Fake(0, 2))
