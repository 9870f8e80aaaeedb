#!/usr/bin/env python
"""
bench.py -- the FFTPower hot path on synthetic log-normal particles.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full `FFTPower(cat, mode='1d', Nmesh=...)` call: paint (CIC, f8 mesh) -> r2c ->
compensate -> |delta(k)|^2 V -> P(k) shell binning -> BinnedStatistic on the host.

Workload (BASELINE.json configs[1], "C2"): ~1e8 float32 log-normal particles in a 1024 Mpc/h box ->
512^3 mesh per GPU.  N > 1 is WEAK scaling: every rank contributes an independent ~1e8-particle
log-normal tile and the global mesh doubles along x, y, z in turn (N=8: 1024^3, ~8e8 particles --
the size of configs[3]); particles are routed to x-slab owners, the FFT does its NCCL all-to-all,
the histogram is all-reduced.

Prints ONE JSON line (rank 0).  `value` = particles/s through the whole step with columns resident
in HBM; `e2e` = the same call fed from pinned HOST arrays (H2D inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMESH_1GPU = 512
BOX_1GPU = 1024.0
NPART_1GPU = 1.0e8
GEN_NMESH = 256


def mesh_for(ngpu):
    """global (Nmesh, BoxSize, tile grid) for the weak-scaling family"""
    f = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[ngpu]
    return [NMESH_1GPU * a for a in f], [BOX_1GPU * a for a in f], f


def ncu_traffic(kernels):
    """DRAM bytes (read + write) per launch of the named kernels, summed, from the newest committed ncu launch list
    under profiles/ (`--metrics ...,dram__bytes_read.sum,dram__bytes_write.sum` pass of this same command); None if
    no list holds them.  Also returns the file used."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launches_bench_*.csv")), reverse=True):
        try:
            rows = [r for r in csv.reader(open(path)) if len(r) > 10]
            hdr = rows[0]
            iname, imet, ival, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
        except Exception:
            continue
        seen, total = {}, 0.0
        for r in rows[1:]:
            if not r[imet].startswith("dram__bytes_"):
                continue
            for k in kernels:
                if k in r[iname] and seen.setdefault(k, r[iid]) == r[iid]:     # first launch of each kernel only
                    total += float(r[ival].replace(",", ""))
        if len(seen) == len(kernels):
            return total, os.path.relpath(path, ROOT)
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows]
        pw = [float(r[3]) for r in rows if r[3].strip().replace(".", "").isdigit()]
        # samples under load: upper half of the power draw
        if pw:
            thr = 0.5 * (max(pw) + min(pw))
            load = [s for s, r in zip(sm, rows) if float(r[3]) >= thr] or sm
        else:
            load = sm
        out["sm_mhz"] = float(np.median(load))
        out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for j, n in enumerate(names):
            if any("Active" == r[5 + j].strip() for r in rows):
                out["reasons"].append(n)
        out["samples"] = len(rows)
        return out


def generate_tile(seed):
    """one ~1e8-particle log-normal tile on the current GPU (device float32 positions, cell-sorted like the
    reference's generator output)"""
    from nbodykit_b200.comm import SelfComm
    from nbodykit_b200.cosmology import NoWiggleEHPower
    from nbodykit_b200.source.catalog.lognormal import LogNormalCatalog
    nbar = NPART_1GPU / BOX_1GPU ** 3
    cat = LogNormalCatalog(Plin=NoWiggleEHPower(), nbar=nbar, BoxSize=BOX_1GPU, Nmesh=GEN_NMESH, bias=2.0, seed=seed,
                           comm=SelfComm())
    return cat['Position'].compute()


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    # keep stdout clean for the ONE JSON line: libraries (e.g. the NCCL version banner) write to fd 1 directly
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nbodykit_b200 import CurrentMPIComm, _lib
    from nbodykit_b200.lab import ArrayCatalog, FFTPower
    comm = CurrentMPIComm.get()
    assert comm.size == world
    Nmesh, Box, tiles = mesh_for(world)

    pos = generate_tile(seed=42 + rank)
    # place this rank's ~1e8 particles in its own x-slab of the global box (an affine stretch of the unit tile: same
    # particles per mesh cell everywhere).  Slab-local input is what the reference's own generators emit
    # ("exchange ~ ghosts only"); decompose / ghost routing still run every step.
    if world > 1:
        slab_w = Box[0] / world
        pos[:, 0] *= slab_w / BOX_1GPU
        pos[:, 0] += slab_w * rank
        for d in (1, 2):
            if Box[d] != BOX_1GPU:
                pos[:, d] *= Box[d] / BOX_1GPU
    n_local = int(pos.shape[0])
    n_total = int(comm.allreduce(n_local))
    cat = ArrayCatalog({'Position': pos}, comm=comm, BoxSize=Box)

    def step(c):
        return FFTPower(c, mode='1d', Nmesh=Nmesh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler needs ~1 s to come up: start it before the warm-up; "under load" samples are picked by
    # power draw when the numbers are reduced
    sampler = ClockSampler(local) if rank == 0 else None
    t_w = time.time()
    nw = 0
    while nw < args.warmup or (time.time() - t_w < 1.5 and nw < 200):
        r = step(cat)
        nw += 1
    barrier()

    l0 = _lib.launch_count()
    _lib.profiler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        r = step(cat)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    stages = _lib.profiler.stop()
    if _lib.profiler.host and rank == 0:
        for k, (t, c) in sorted(_lib.profiler.wall.items()):
            sys.stderr.write("TRACE %-20s %8.3f ms/step (%d calls)\n" % (k, 1e3 * t / max(1, nw + args.steps), c))
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps

    # ---- e2e: same call, columns start in pinned host memory
    host = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=True)
    host.copy_(pos)
    del cat
    torch.cuda.synchronize()

    def step_host():
        c = ArrayCatalog({'Position': host}, comm=comm, BoxSize=Box)
        return FFTPower(c, mode='1d', Nmesh=Nmesh)

    del pos
    step_host()
    barrier()
    e0.record()
    for _ in range(args.steps):
        rh = step_host()
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_e2e = float(tms.item()) / args.steps
    d2h = int(r.power.data.nbytes) + 8 * 3 * (len(r.power['k']) + 2) * 3   # packed histogram read back

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, which = peaks()
    paint_ms = float(np.mean(stages.get("paint", [float('nan')])))
    mesh_cells = float(np.prod(Nmesh)) / world
    alg_bytes = n_local * 12.0 + mesh_cells * 8.0          # particles read once + mesh written once
    achieved = alg_bytes / (paint_ms * 1e-3) / 1e9
    stage_ms = {k: float(np.mean(v)) for k, v in stages.items()}
    traffic, traffic_src = (ncu_traffic(["k_tile_count_blk", "k_tile_colscan", "k_tile_scan", "k_tile_scatter_blk",
                                         "k_tile_paint"]) if world == 1 else (None, None))
    # the other two stages against the same peak (algorithmic bytes of SURVEY.md 8d: 4 x field for the 3-D r2c,
    # one read of the complex field for the fused |delta_k|^2 binning)
    other = {}
    field_bytes = mesh_cells * 8.0
    cplx_bytes = mesh_cells / Nmesh[2] * (Nmesh[2] // 2 + 1) * 16.0
    if world == 1 and "r2c" in stage_ms:
        a = 4.0 * field_bytes / (stage_ms["r2c"] * 1e-3) / 1e9
        other["r2c"] = {"algorithmic_bytes": 4.0 * field_bytes, "kernel_ms": stage_ms["r2c"], "achieved": a, "frac": a / hbm}
    if "power_bin" in stage_ms:
        a = cplx_bytes / (stage_ms["power_bin"] * 1e-3) / 1e9
        other["power_bin"] = {"algorithmic_bytes": cplx_bytes, "kernel_ms": stage_ms["power_bin"], "achieved": a, "frac": a / hbm}
    out = {
        "metric": "particles/sec painted + P(k) end-to-end (FFTPower 1d, CIC, f8 mesh)",
        "value": n_total / (ms_step * 1e-3),
        "unit": "particles/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "LogNormal %.3g particles (f4) -> %s mesh CIC f8 compensated, FFTPower mode=1d (BASELINE configs[1] per GPU)"
                               % (n_total, "x".join(str(v) for v in Nmesh)),
                   "particles": n_total, "Nmesh": Nmesh, "BoxSize": Box, "resampler": "cic", "mesh_dtype": "f8",
                   "l2": "inputs (1.2 GB particles, 1.07 GB mesh per GPU) exceed the 126 MB L2; no flush needed",
                   "parallelism": "x-slab x%d" % world,
                   "placement": "each rank's particles lie in its own x-slab (slab-local generator output); "
                                "decompose + ghost exchange run inside every step"},
        "paint_particles_per_sec": n_total / (paint_ms * 1e-3),
        "pk_seconds": ms_step * 1e-3,
        "stage_ms": stage_ms,
        "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": "particles/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": n_local * 12 * world, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "paint = nbk_paint_tiled (k_tile_count_blk, k_tile_colscan, k_tile_scan, k_tile_scatter_blk, "
                               "k_tile_paint; the mesh clear rides in the count pass)",
                     "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "peak_source": which, "algorithmic_bytes": alg_bytes,
                     "kernel_ms": paint_ms, "traffic": traffic, "traffic_source": traffic_src,
                     "other_stages": other},
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(host.numpy(), Nmesh, Box)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(pos, Nmesh, Box, sample=2 * 10 ** 7, mesh_seconds=None):
    """the CPU oracle (port of the reference flow) on the host cores: paint rate from a bounded sample of the
    particles, mesh stages (r2c, compensate, |delta|^2, binning) at full mesh size"""
    from oracle import build_c, pmesh_oracle as po
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    n = len(pos)
    ns = min(sample, n)
    t0 = time.time()
    mesh = build_c.paint(pos[:ns], None, Nmesh, Box, "cic")
    t_paint = time.time() - t0
    if mesh_seconds is None:
        t0 = time.time()
        mesh /= (ns / float(np.prod(Nmesh)))
        c = po.r2c(mesh)
        del mesh
        c = po.compensate("CompensateCICShotnoise", po.k_coords(Nmesh, Box, "f4", kind="circular"), c)
        res = po.power_from_complex(c, None, Nmesh, Box, mode="1d")
        t_mesh = time.time() - t0
    else:
        t_mesh = mesh_seconds      # the mesh stages do not depend on the particle sample: re-used (see `sample`)
    total = n * (t_paint / ns) + t_mesh
    return {"value": n / total, "unit": "particles/s", "cores": cores, "kind": "port",
            "paint_particles_per_sec": ns / t_paint, "mesh_seconds": t_mesh,
            "sample": "paint: first %d of %d particles (C/OpenMP restatement of the pmesh scatter, %d threads), "
                      "extrapolated linearly; r2c (scipy pocketfft, %d workers) + compensate + project_to_basis "
                      "(NumPy restatement) at the full %s mesh" % (ns, n, cores, cores, "x".join(str(v) for v in Nmesh))}


def run_reference(args):
    """the reference's CPU algorithm for this path (oracle port: /root/reference needs pmesh, absent here),
    all host threads, on a bounded sample of the same workload"""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    Nmesh, Box, _ = mesh_for(1)
    # same generator as our arm when a GPU is there; else a uniform stand-in of the same size
    if torch.cuda.is_available():
        torch.cuda.set_device(0)
        pos = generate_tile(seed=42).cpu().numpy()
    else:
        pos = (np.random.RandomState(42).uniform(size=(int(NPART_1GPU), 3)) * BOX_1GPU).astype("f4")
    vals = []
    base = None
    t_start = time.time()
    budget = 240.0                  # seconds for the whole --warmup W --steps K run
    mesh_times = []
    reused = 0
    for i in range(args.warmup + args.steps):
        left = args.warmup + args.steps - i
        # every step times the particle sample; the full-mesh stages (r2c + compensate + binning, independent of the
        # sample) are re-timed as long as the remaining steps fit the budget, otherwise their mean so far is re-used
        reuse = None
        if mesh_times and (time.time() - t_start) + left * (np.mean(mesh_times) + 1.0) > budget:
            reuse = float(np.mean(mesh_times))
            reused += 1
        base = cpu_baseline(pos, Nmesh, Box, sample=10 ** 7, mesh_seconds=reuse)
        if reuse is None:
            mesh_times.append(base["mesh_seconds"])
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    if reused:
        base["sample"] += "; mesh stages timed in %d of %d steps (time budget %.0f s), their mean re-used in the rest" % (
            len(mesh_times), args.warmup + args.steps, budget)
    n = len(pos)
    out = {"impl": "reference", "metric": "particles/sec painted + P(k) end-to-end (FFTPower 1d, CIC, f8 mesh)",
           "value": v, "unit": "particles/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": n / v * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "LogNormal %.3g particles (f4) -> 512x512x512 mesh CIC f8 compensated, FFTPower mode=1d" % n,
                      "particles": n, "Nmesh": Nmesh, "BoxSize": Box},
           "cpu_baseline": base,
           "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
