#!/usr/bin/env python
"""bench.py - Lanczos steps/s and achieved HBM GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4|c5|c3full] [--impl reference]

One "step" = one Lanczos step (operator apply + alpha, three-term update + beta, and the
re-orthogonalisation sweeps the configuration asks for).  The default workload is BASELINE
config 3 - the 512^3 periodic 7-point Laplacian (134 M unknowns, fp64, selective
re-orthogonalisation) on which the north-star roofline target is stated; with N > 1 GPUs the
grid is weak-scaled to 512 x 512 x (512 N), row-sharded in z-slabs (one process per GPU).

Prints ONE JSON line (rank 0).  `value` is device-timed with the start vector resident in HBM;
`e2e` goes through the drop-in class with a pinned HOST start vector and host results.
`--impl reference` times the oracle port of the reference's CPU path on this box's host cores.
"""
from __future__ import annotations

import argparse
import builtins
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c3": dict(desc="3D 7-point periodic Laplacian 512x512x(512*gpus) fp64, selective reorth (CGS2 when triggered)",
               grid=(512, 512, 512), steps=100, reorth="selective", cgs_passes=2),
    "c3full": dict(desc="3D 7-point periodic Laplacian 512^3 fp64, full reorth (reference form)",
                   grid=(512, 512, 512), steps=60, reorth="full", cgs_passes=1),
    "c2": dict(desc="graph Laplacian of a 2D Delaunay mesh, 1M vertices, SELL-32-2048, full reorth",
               npts=1_000_000, steps=200, reorth="full", cgs_passes=1),
    "c4": dict(desc="graph Laplacian of a 3D random geometric graph (Poisson points, mean degree 13, ~14 nnz/row), "
                    "~50M vertices in cell order, SELL-32-2048, selective reorth, z-slabs of cells over the GPUs (strong scaling)",
               cells=(253, 253, 252), steps=100, reorth="selective", cgs_passes=2, strong=True),
    "c5": dict(desc="3D 7-point periodic Laplacian 1024^3 (1.07 B unknowns) fp64, CGS2 every step, z-slabs over the GPUs (strong scaling)",
               grid=(1024, 1024, 1024), steps=60, reorth="full", cgs_passes=2, strong=True),
    "c1": dict(desc="2D 5-point Dirichlet Laplacian 200x200 fp64, full reorth",
               grid=(200, 200), steps=100, reorth="full", cgs_passes=1),
    # the reference's own production size (3Ddeuteron.py:63-64): N = 160, n = 400, 27-point T, H = -T + V, built
    # by the drop-in Hamiltonian class (potential traced from the NumPy function and evaluated on the device)
    "deut": dict(desc="3Ddeuteron.py: 27-point H = -T + V on 160^3 (4.1 M unknowns), n = 400, full reorth (reference form)",
                 grid=(160, 160, 160), steps=400, reorth="full", cgs_passes=1, deuteron=True),
}
METRIC = "lanczos_steps_per_sec"
UNIT = "steps/s"


def quiet_banners():
    """The drop-in classes print the reference's '+++' banners; keep stdout to the JSON line."""
    real = builtins.print

    def p(*a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("+++"):
            return
        real(*a, **k)
    builtins.print = p


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons, pw = [], [], set(), []
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, cmax = float(f[0]), float(f[1])
            except ValueError:
                continue
            inside = (t0 - 0.05) <= t <= (t1 + 0.15)
            if inside:
                sm.append(clk)
                try:
                    pw.append(float(f[2]))
                except ValueError:
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            mx.append(cmax)
        if not sm:   # region shorter than the sampling period: use every sample
            for t, line in self.lines:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel_name):
    """(DRAM bytes per launch, source file) of `kernel_name`: dram__bytes_read.sum + dram__bytes_write.sum from the
    committed `ncu --set full` summaries under profiles/ (newest round first; captured at the 512^3 config / the
    50M-vertex graph).  Static evidence, not measured in this run - the source file is named in the JSON line."""
    for name in ("r2_ncu_traffic.json", "r1c_ncu_traffic.json", "r1b_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            for k, v in t.items():
                if k.startswith(kernel_name):
                    return v["dram_bytes_per_launch"], "profiles/" + name
        except Exception:
            pass
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- CPU baseline
def _quiet_call(fn, *a, **k):
    """The reference prints banners and draws tqdm bars (Lanczos.py:79,111): keep stdout/stderr clean."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def _reference_runner(irregular: bool):
    """(run(H, n, v0, sweeps) -> seconds, kind).  `kind` "reference": the unmodified reference classes
    vendored by oracle/build_ref.py into oracle/_ref (Lanczos.execute_Lanczos / IrrLanczos.execute_LanczosOld,
    use_cuda=False); "port": the oracle's restatement of the same NumPy/SciPy operations when oracle/_ref is
    absent.  sweeps=False times the same loop with the reference's `reorthogonalize` call made a no-op - the
    like-for-like of a GPU run in which no Gram-Schmidt sweep fired (the reference has no such mode)."""
    from oracle import build_ref
    from oracle import lanczos_oracle as orc
    mods = build_ref.load()
    if mods is None:
        def run_port(H, n, v0, sweeps=True):
            return orc.timed_steps(H, n, v0, reorth=sweeps)
        return run_port, "port"
    cls = mods[1].IrrLanczos if irregular else mods[0].Lanczos

    def run_ref(H, n, v0, sweeps=True):
        saved = cls.__dict__["reorthogonalize"]
        if not sweeps:
            cls.reorthogonalize = staticmethod(lambda V, j, use_cuda=True: None)
        try:
            L = cls(H)
            t0 = time.perf_counter()
            _quiet_call(L.execute_LanczosOld if irregular else L.execute_Lanczos, n, use_cuda=False, v0=v0)
            return time.perf_counter() - t0
        finally:
            cls.reorthogonalize = saved
    return run_ref, "reference"


def cpu_reference_leg(workload: str, budget_s: float = 25.0):
    """The reference's CPU path on a bounded sample of the workload, with all the host threads NumPy/OpenBLAS
    will use (SciPy's SpMV and NumPy's elementwise kernels are single-threaded).  Two numbers:
      reference_form  the reference as it is - one full Gram-Schmidt sweep against all n rows every step;
      no_sweep        the same loop without the sweep, at the largest size that builds in seconds (256^3).
    `value` is the one that matches what the GPU arm does on this workload: no_sweep for the selective /
    none configurations (in the timed region no sweep fires), reference_form for the full-reorth ones.
    Both are extrapolated linearly in the number of unknowns to the workload's size (stated in `sample`)."""
    from oracle import lanczos_oracle as orc
    wl = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    irregular = "grid" not in wl
    run, kind = _reference_runner(irregular)
    out = dict(unit=UNIT, cores=cores, kind=kind, numpy=np.__version__)
    if wl.get("deuteron"):
        full_M, n_full = int(np.prod(wl["grid"])), wl["steps"]
        side, n = 40, 60
        H, *_ = orc.deuteron_hamiltonian(side)
        v0 = np.random.RandomState(99).uniform(-1, 1, side ** 3)
        t = run(H, n, v0, True)
        rate = n / t
        # the sweep against all n rows every step costs ~ n * M per step: extrapolate in both
        value = rate * (side ** 3) / full_M * n / n_full
        out.update(value=value, reference_form_value=value,
                   sample=f"7-point deuteron Hamiltonian at N={side}, n={n}, the reference's full sweep every step: {rate:.2f} steps/s "
                          f"measured, x{side ** 3 / full_M:.2e} (linear in M) x{n / n_full:.2f} (the sweep reads all n rows) -> "
                          f"{value:.4f} steps/s at N=160, n=400")
        return out
    if "grid" in wl and len(wl["grid"]) == 3:
        full_M = int(np.prod(wl["grid"]))
        side, n = 96, 32
        H = orc.laplacian_csr((side,) * 3, 6.0, -1.0, periodic=True)
        v0 = np.random.RandomState(99).uniform(-1, 1, side ** 3)
        t = run(H, n, v0, True)
        form_rate = n / t
        form_value = form_rate * (side ** 3) / full_M
        side2, n2 = 256, 12
        H2 = orc.laplacian_csr((side2,) * 3, 6.0, -1.0, periodic=True)
        v2 = np.random.RandomState(99).uniform(-1, 1, side2 ** 3)
        t2 = run(H2, n2, v2, False)
        ns_rate = n2 / t2
        ns_value = ns_rate * (side2 ** 3) / full_M
        like = wl["reorth"] != "full"
        out.update(value=ns_value if like else form_value, reference_form_value=form_value, no_sweep_value=ns_value,
                   step_only_value=ns_value,
                   sample=(f"reference_form: {side}^3 periodic 7-pt Laplacian (CSR), n={n}, the reference's full sweep every step: "
                           f"{form_rate:.2f} steps/s measured, x{side ** 3 / full_M:.2e} (linear in M) -> {form_value:.4f} steps/s at "
                           f"{full_M} unknowns (its sweep cost also grows with n); no_sweep: {side2}^3, n={n2}, sweep disabled: "
                           f"{ns_rate:.2f} steps/s measured, x{side2 ** 3 / full_M:.3f} -> {ns_value:.3f} steps/s; "
                           f"value = {'no_sweep' if like else 'reference_form'}"))
        return out
    if workload == "c4":
        npts, n = 200_000, 32
        H = orc.rgg_graph_laplacian(npts, mean_degree=13.0, seed=0)
        v0 = np.random.RandomState(99).uniform(-1, 1, npts)
        t = run(H, n, v0, True)
        t2 = run(H, n, v0, False)
        full = 50_000_000
        out.update(value=(n / t2) * npts / full, reference_form_value=(n / t) * npts / full, no_sweep_value=(n / t2) * npts / full,
                   sample=f"3D random geometric graph, {npts} vertices, mean degree 13, n={n}: reference_form {n / t:.2f} steps/s, "
                          f"no_sweep {n / t2:.2f} steps/s measured, both x{npts / full:.1e} (linear in M) to {full} vertices; value = no_sweep")
        return out
    if workload == "c2":
        npts, n = 100_000, 40
        H = orc.delaunay_graph_laplacian(npts, seed=0)
        v0 = np.random.RandomState(99).uniform(-1, 1, npts)
        t = run(H, n, v0, True)
        value = (n / t) * npts / wl["npts"]
        out.update(value=value, reference_form_value=value,
                   sample=f"Delaunay {npts} vertices, n={n}, full sweep every step ({n / t:.2f} steps/s measured), "
                          f"x{npts / wl['npts']:.2f} (linear in M) to {wl['npts']} vertices")
        return out
    grid, n = wl["grid"], wl["steps"]
    H = orc.laplacian_csr(grid, 4.0, -1.0, periodic=False)
    v0 = np.random.RandomState(99).uniform(-1, 1, int(np.prod(grid)))
    t = run(H, n, v0, True)
    out.update(value=n / t, reference_form_value=n / t, sample=f"the full config: {grid}, n={n}, full sweep every step, not extrapolated")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    base = cpu_reference_leg(args.workload)
    v = float(base["value"])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v if v > 0 else None,
            "higher_is_better": True, "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {wl['desc']}",
                                            "note": "one bounded sample per run (see cpu_baseline.sample); --steps/--warmup do not "
                                                    "change it: the reference at the full size would take hours"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- parity gate
def parity_check(lz, world, rank, dist):
    """Before anything is timed: a small solve through the code path this run uses - row-sharded over the
    `world` GPUs of the run (a 2-shard single-process team on one GPU at N = 1, so that the peer-memory
    exchange is always exercised) - held against the oracle and against the single-GPU result, alpha/beta
    <= 1e-12 relative (n = 40 <= 50, full re-orthogonalisation in the reference's form).  The oracle is used
    as the checker only.  Returns the dict that goes into the JSON line; `ok` False makes bench.py exit 1."""
    import torch
    from lanczos_b200 import team as lzteam
    from oracle import lanczos_oracle as orc
    n, tol = 40, 1e-12
    out = {"tol": tol, "n": n, "cases": []}

    def rel(a, b):
        return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))

    def record(name, Hs, ref, single, sharded):
        a1, b1 = np.diag(single.H_eff), np.diag(single.H_eff, 1)
        a2, b2 = np.diag(sharded.H_eff), np.diag(sharded.H_eff, 1)
        c = {"case": name, "shards": sharded.world,
             "sharded_vs_oracle": max(rel(a2, ref["alpha"]), rel(b2, ref["beta"])),
             "single_vs_oracle": max(rel(a1, ref["alpha"]), rel(b1, ref["beta"])),
             "sharded_vs_single": max(rel(a2, a1), rel(b2, b1))}
        c["ok"] = bool(max(c["sharded_vs_oracle"], c["single_vs_oracle"], c["sharded_vs_single"]) < tol)
        out["cases"].append(c)

    shards = world if world > 1 else 2
    # structured grid, z-slabs + halo planes
    grid = (64, 64, 16 * shards)
    H = orc.laplacian_csr(grid, 6.0, -1.0, periodic=True)
    ref = orc.lanczos(H, n, seed=7)
    op = lz.StencilOperator(grid, 6.0, -1.0)
    single = lz.Lanczos(op)
    single.execute_Lanczos(n, seed=7, use_cuda=False, verbose=False)
    sharded = lzteam.TeamLanczos(op, rank=rank, world=world) if world > 1 else lzteam.LocalTeamLanczos(op, shards)
    sharded.execute_Lanczos(n, seed=7)
    record("7-point periodic %dx%dx%d, z-slabs" % grid, H, ref, single, sharded)
    del sharded
    # sparse rows, ghost-index exchange
    G = orc.delaunay_graph_laplacian(4000 * shards, seed=1)
    refg = orc.lanczos(G, n, seed=7)
    singleg = lz.IrrLanczos(G)
    singleg.execute_LanczosOld(n, seed=7, verbose=False)
    shardedg = lzteam.TeamLanczos(G, rank=rank, world=world) if world > 1 else lzteam.LocalTeamLanczos(G, shards)
    shardedg.execute_Lanczos(n, seed=7)
    record("Delaunay graph Laplacian %d vertices, SELL row blocks" % G.shape[0], G, refg, singleg, shardedg)
    del shardedg
    torch.cuda.synchronize()
    ok = all(c["ok"] for c in out["cases"])
    if world > 1:                       # every rank must agree that it passed
        t = torch.tensor([0 if ok else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(t)
        ok = int(t.item()) == 0
    out["ok"] = bool(ok)
    out["mode"] = ("one process per GPU over %d GPUs (TeamLanczos)" % world) if world > 1 else \
                  "2 shards driven by one process on the one GPU (LocalTeamLanczos)"
    return out


# --------------------------------------------------------------------------- GPU arm
def build_operator(lz, workload, world, rank):
    wl = WORKLOADS[workload]
    if wl.get("deuteron"):
        N, L = wl["grid"][0], 25

        def potential(x, y, z):                              # 3Ddeuteron.py:51-61
            r = np.sqrt(x**2 + y**2 + z**2)
            eWell = 54.531
            eWells = 65.4823128982115
            eCores = 40.0*eWell
            rCore = 1.0/4
            rWell = 17.0/10
            fPow = 4.0
            return eCores*np.exp(-(r/rCore)**fPow) - eWells*np.exp(-(r/rWell)**fPow)
        dx = float(L) / N
        T_factor = 197.327**2/(2*469.4592) * 1/dx**2         # 3Ddeuteron.py:67-70
        system = lz.Hamiltonian(N, L, potential, T_factor)
        system.create_sparse_T()
        system.create_sparse_V()
        H = (-system.T_sparse + system.V_sparse)
        H.sort_indices()
        return H, tuple(wl["grid"])
    if "grid" in wl:
        grid = tuple(wl["grid"])
        dim = len(grid)
        if dim == 3 and world > 1 and not wl.get("strong"):
            grid = (grid[0], grid[1], grid[2] * world)       # weak scaling in z
        return lz.StencilOperator(grid, 2.0 * dim, -1.0, bc="periodic" if dim == 3 else "dirichlet"), grid
    if "cells" in wl:
        # generated on the device, one row block per GPU (include/lz_synth.h)
        from lanczos_b200 import synth
        from lanczos_b200.engine import DeviceCSR
        gen = synth.RggGenerator(wl["cells"], seed=0)
        if world > 1:
            return gen.row_block(rank, world), (gen.M,)
        return DeviceCSR(*gen.rows(0, gen.M)), (gen.M,)
    from oracle import lanczos_oracle as orc             # input generator only (test infrastructure)
    H = orc.delaunay_graph_laplacian(wl["npts"], seed=0)
    return H, (wl["npts"],)


def bind_to_gpu_numa_node(local):
    """Pin this process (and the pinned host buffers it first-touches afterwards) to the CPUs of the NUMA node
    the GPU hangs off: with 4-8 ranks staging 1 GB start vectors at once, remote-node pinned memory is what
    makes the end-to-end number fall off.  Best effort; returns a description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return "single NUMA node"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return "node %d has no allowed CPUs" % node
        os.sched_setaffinity(0, allowed)
        return "bound to NUMA node %d (%d CPUs)" % (node, len(allowed))
    except Exception as e:
        return "not bound (%s)" % type(e).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-cgs-fusion", action="store_true", help="CGS2 as four separate sweeps (comparison runs)")
    ap.add_argument("--kb-alpha", action="store_true",
                    help="recompute step with alpha accumulated inside KB + border kernel instead of a KA pass (comparison runs)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="sparse row shards without the interior/boundary overlap on a second stream (comparison runs)")
    ap.add_argument("--min-region-s", type=float, default=1.2,
                    help="the K-step solve is repeated until the timed region is at least this long")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "two_pass", "recompute", "fused"],
                    help="auto = the library default (recompute for matrix-free operators)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = wl["steps"]
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    quiet_banners()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator comes up; stdout carries
        # exactly one JSON line, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    numa = bind_to_gpu_numa_node(local)
    import lanczos_b200 as lz

    # ---- parity gate: nothing is timed unless the code path of this run agrees with the oracle ----------------
    parity = None
    if not args.no_parity_check:
        parity = parity_check(lz, world, rank, dist)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity_check": parity,
                                  "error": "parity check failed: nothing was timed"}), flush=True)
            raise SystemExit(1)

    K, W = int(args.steps), max(3, int(args.warmup))
    H, grid = build_operator(lz, args.workload, world, rank)
    M_total = int(np.prod(grid))
    opts = dict(reorth=wl["reorth"], cgs_passes=wl["cgs_passes"], ref_compat=True, step_kernel=args.step_kernel,
                cgs_fused=not args.no_cgs_fusion, kb_alpha=args.kb_alpha)

    if world > 1:
        from lanczos_b200 import team as lzteam
        solver = lzteam.TeamLanczos(H, rank=rank, world=world)
        M_local = solver.M_local
        opts["overlap"] = not args.no_overlap
    else:
        solver = lz.Lanczos(H) if "grid" in wl else lz.IrrLanczos(H)
        M_local = M_total
    execute = solver.execute_Lanczos if hasattr(solver, "execute_Lanczos") and "grid" in wl else solver.execute_LanczosOld

    # start vector: device-generated for the throughput run, pinned host copy for e2e
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    v0_dev = torch.rand(M_local, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v0_host = torch.empty(M_local, dtype=torch.float64).pin_memory()
    v0_host.copy_(v0_dev)
    torch.cuda.synchronize()

    # basis memory: K rows of 8*M bytes must fit; otherwise split the K steps into several solves
    free_b, _ = torch.cuda.mem_get_info()
    row_b = 8 * ((M_local + 63) // 64 * 64)
    max_rows = max(2, int((free_b * 0.85 - 5 * row_b) // row_b))
    if wl["reorth"] == "none":
        max_rows = K
    chunks = []
    left = K
    while left > 0:
        c = min(left, max_rows)
        if left - c == 1:          # a 1-step solve is not allowed in ref_compat (n >= 2)
            c -= 1
        chunks.append(c)
        left -= c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(n, v0, profile=False):
        # a 1-step solve (--steps 1) cannot follow the reference loop, which needs n >= 2 (Lanczos.py:107,112)
        execute(n, v0=v0, profile=profile, **dict(opts, ref_compat=opts["ref_compat"] and n >= 2))
        return solver.result

    # ---- warm-up: W untimed steps (also sizes the workspace arena and the allocator cache) ----
    solve(max(W, 2), v0_dev)
    res = None
    for c in chunks:
        res = solve(c, v0_dev)             # allocator + arena warm at the timed size
    est_ms = max(1e-3, sum(1 for _ in chunks) * 0 + res.gpu_ms * K / chunks[-1])
    del res
    if world > 1:
        t = torch.tensor([est_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        est_ms = float(t.item())
    # the K-step solve is repeated R times so that the timed region lasts >= min_region_s: long enough for
    # >= 10 clock samples and for the board to reach its sustained power state
    R = int(min(2000, max(1, np.ceil(args.min_region_s * 1e3 / est_ms))))
    barrier()

    def profile_pass(nsolves):
        """K steps `nsolves` times with a CUDA-event pair around every bandwidth kernel -> {kind: [ms, launches]}."""
        kern_, sk = {}, "two_pass"
        for _ in range(nsolves):
            for c in chunks:
                res_ = solve(c, v0_dev, profile=True)
                sk = res_.step_kernel
                for k_, (ms_, cnt_) in res_.kernel_ms.items():
                    acc_ = kern_.setdefault(k_, [0.0, 0])
                    acc_[0] += ms_
                    acc_[1] += cnt_
                del res_
        for k_ in ("apply", "update", "dots", "gs_update", "fused", "gs_fused", "border"):
            kern_.setdefault(k_, [0.0, 0])
        return kern_, sk

    # ---- region A0: per-kernel timings in BURST conditions - before the long region has driven the board into its
    # power cap - because the roofline's denominator (MEASURED_PEAKS.json hbm_gbs) is a burst figure too ("best of 10")
    prof_solves = max(1, min(R, 3))
    kern, step_kernel = profile_pass(prof_solves)
    barrier()
    time.sleep(0.5)

    # ---- timed region A: device-resident input, R x exactly K steps ----------------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    launches = 0
    reorths = 0
    solve_ms = []
    alpha_in_update = False
    overlap_used = False
    for _ in range(R):
        ms = 0.0
        for c in chunks:
            res = solve(c, v0_dev)
            launches += res.launches
            reorths += res.reorth_count
            ms += res.gpu_ms
            alpha_in_update = getattr(res, "alpha_in_update", False)
            overlap_used = getattr(res, "overlap", False)
            del res                       # the result owns the basis (~100 GB at K = 100); the next solve reuses it
        solve_ms.append(ms)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_dev = ev0.elapsed_time(ev1)
    burst_ms = min(solve_ms)
    if world > 1:
        t = torch.tensor([ms_dev, burst_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, burst_ms = float(t[0].item()), float(t[1].item())
    time.sleep(0.1)
    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- region A2: the same per-kernel timings right after the long region (sustained conditions: the board sits at
    # its power cap, clocks are down).  Kept out of region A because the event records break up back-to-back launches.
    kern_hot, _ = profile_pass(prof_solves)
    barrier()
    theta = solver.ritz_values(10)

    # ---- timed region B: end to end through the drop-in class, host buffers --------------------
    # every solve copies its start vector from pinned host memory and returns alpha/beta to the host
    barrier()
    e0 = time.perf_counter()
    for c in chunks:
        solve(c, v0_host)
    torch.cuda.synchronize()
    e_one = time.perf_counter() - e0
    if world > 1:
        t = torch.tensor([e_one], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_one = float(t.item())
    Re = int(min(500, max(2, np.ceil(min(args.min_region_s, 1.0) / max(e_one, 1e-4)))))
    barrier()
    e0 = time.perf_counter()
    for _ in range(Re):
        for c in chunks:
            solve(c, v0_host)
            T = solver.H_eff                   # host ndarray (alpha/beta came back D2H inside the call)
            _ = solver.ritz_values(10)
    torch.cuda.synchronize()
    e_wall = time.perf_counter() - e0
    if world > 1:
        t = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_wall = float(t.item())

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = measured_peaks()
    N = M_local
    is_stencil = "grid" in wl
    recompute = step_kernel == "recompute"
    if is_stencil:
        # two-pass: K1 reads v_j, writes w (alpha fused); recompute: KA reads v_j only
        apply_bytes = 8.0 * N if recompute else 16.0 * N
    else:
        nnz_true, nnz_stored = solver.nnz_local() if world > 1 else solver._device_op.nnz()
        value_free = solver.value_free_local() if world > 1 else solver._device_op.value_free()
        # 12 B per entry (value + column index) + x and y.  An unweighted graph Laplacian is applied
        # from its column indices alone: 4 B per entry + x, y and the per-row diagonal coefficient
        # windowed form (csrc/sellw.cu): 16-bit offsets into a shared-memory stage of x, 2 B less per entry
        windowed = solver.windowed_local() if world > 1 else solver._device_op.windowed()
        idx_bytes = 2.0 if windowed else 4.0
        apply_bytes = (idx_bytes * nnz_true + 24.0 * N) if value_free else (12.0 * nnz_true + 16.0 * N)
    per_kernel = {}
    # K3 reads w, v_j, v_{j-1}, writes r; KB (recompute) reads v_j, v_{j-1}, writes r and applies H again
    update_bytes = 24.0 * N if recompute else 32.0 * N
    # border kernel (alpha inside KB): 2 of every 8 rows (y tile borders), the two sectors around every 64th
    # column (x tile borders, 64 B per 512 B of row) and two planes per z-chunk - counted as 3.5 B per unknown
    border_bytes = 3.5 * N
    alg = {"apply": apply_bytes, "update": update_bytes, "border": border_bytes}
    for k in ("apply", "update", "border"):
        ms, cnt = kern[k]
        if cnt:
            per_kernel[k] = {"launches": cnt, "avg_ms": ms / cnt, "alg_bytes": alg[k],
                             "achieved_gbs": alg[k] / (ms / cnt) / 1e6}
    # Gram-Schmidt kernels: the bytes depend on the row count of each launch; with full
    # re-orthogonalisation the schedule is known: at step j every sweep reads the j rows before row j
    # and row j itself ((2k+3)*8*N bytes per sweep against k vectors).
    if wl["reorth"] == "full":
        dots_b = upd_b = fus_b = 0.0
        gs_fused = kern["gs_fused"][1] > 0                    # CGS2: K4c = update of sweep 1 + dots of sweep 2
        for c in chunks:
            for j in range(c):
                for p_ in range(wl["cgs_passes"]):
                    if gs_fused and j >= 1 and p_ == 0:
                        dots_b += (j + 1) * 8.0 * N
                        fus_b += (j + 2) * 8.0 * N            # j rows + target in, target out, dots of sweep 2 for free
                    elif gs_fused and j >= 1 and p_ == 1:
                        upd_b += (j + 2) * 8.0 * N
                    else:
                        dots_b += (j + 1) * 8.0 * N           # j basis rows + the target row
                        upd_b += (j + 2) * 8.0 * N            # j rows + target in, target out
        for k, tot_b in (("dots", dots_b), ("gs_update", upd_b), ("gs_fused", fus_b)):
            ms, cnt = kern[k]
            if cnt:
                per_kernel[k] = {"launches": cnt, "avg_ms": ms / cnt, "alg_bytes": tot_b * prof_solves / cnt,
                                 "achieved_gbs": tot_b * prof_solves / ms / 1e6}
    gs_ms = (kern["dots"][0] + kern["gs_update"][0] + kern["gs_fused"][0]) / prof_solves
    dom = max(per_kernel, key=lambda k: per_kernel[k]["avg_ms"] * per_kernel[k]["launches"]) if per_kernel else None
    names = {"apply": "stencil_alpha_kernel (KA2)" if recompute else ("stencil_apply_dot_kernel" if is_stencil else "spmv_sell_dot_kernel"),
             "update": ("stencil_apply_dot_kernel<MODE=2, ALPHA> (KB + alpha of the next vector)" if alpha_in_update else
                        "stencil_apply_dot_kernel<MODE=2> (KB)") if recompute else "update_norm_kernel",
             "border": "stencil_alpha_border_kernel",
             "dots": "cgs_dots_kernel", "gs_update": "cgs_update_kernel", "gs_fused": "cgs_update_dots_kernel"}
    ncu_names = {"apply": "stencil_alpha_fast_kernel<0>" if recompute else names["apply"],
                 "update": ("stencil_apply_dot_kernel<2, 1, 1, 0, 2, 1>" if alpha_in_update else
                            "stencil_apply_dot_kernel<2, 1, 1, 0, 2, 0>") if recompute else names["update"],
                 "border": names["border"],
                 "dots": names["dots"], "gs_update": names["gs_update"], "gs_fused": names["gs_fused"]}
    roofline = None
    if dom:
        pk = per_kernel[dom]
        traffic, traffic_src = (None, None)
        if (args.workload in ("c3", "c3full") and M_local == 512 ** 3) or (args.workload == "c4" and world == 1):
            traffic, traffic_src = ncu_traffic(ncu_names[dom])
        roofline = {"kernel": names[dom], "bound": "hbm", "achieved": pk["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": pk["achieved_gbs"] / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src, "alg_bytes_per_launch": pk["alg_bytes"],
                    "avg_launch_ms": pk["avg_ms"], "launches": pk["launches"],
                    "timing": "CUDA-event pair around every launch of %d K-step solves on the launching stream, taken BEFORE the "
                              "long timed region (burst conditions, like the peak it is divided by); the same kernels timed right "
                              "after the region are in kernels_sustained" % prof_solves,
                    "frac_of_8TBs_nominal": pk["achieved_gbs"] / 8000.0}
    ms_per_step = ms_dev / (R * K)
    # whole-job aggregate: every rank advances its own 512^3-unknown shard K steps (weak scaling),
    # so the job processes world*K shard-steps; at N = 1 this is plain Lanczos steps/s.
    strong = bool(wl.get("strong"))
    scale_n = 1 if strong else world
    value = scale_n * R * K / (ms_dev / 1e3)
    plain = reorths == 0
    if recompute and alpha_in_update:
        step_bytes = update_bytes + border_bytes
        step_note = ("bytes per plain step: 24*N (KB: H v re-evaluated inside the three-term update, alpha of the new vector "
                     "accumulated while it is in registers) + ~3.5*N (border kernel: the edges that cross CTA tiles)")
    else:
        step_bytes = apply_bytes + update_bytes
        step_note = ("bytes per plain step: 32*N with the recompute step (KA2 8N + KB 24N: H v is re-evaluated instead of "
                     "written and re-read), 48*N with the two-pass step (SURVEY 8d), + the operator's own bytes for stored operators")
    moved_gbs = step_bytes / ms_per_step / 1e6
    fused = {"step_kernel": step_kernel, "alpha_in_update": bool(alpha_in_update), "overlap": bool(overlap_used),
             "moved_bytes_per_step": step_bytes,                       # what this implementation actually moves
             "achieved_gbs": moved_gbs if plain else None,
             "frac_of_measured_peak": moved_gbs / peak if plain else None,
             "frac_of_8TBs_nominal": moved_gbs / 8000.0 if plain else None,
             "burst_achieved_gbs": step_bytes / (burst_ms / K) / 1e6 if plain else None,
             "burst_frac_of_measured_peak": step_bytes / (burst_ms / K) / 1e6 / peak if plain else None,
             "note": step_note + "; sustained over the whole timed region, burst_* = the fastest single K-step solve "
                                 "(both include the pre-step, amortised over K); null when Gram-Schmidt sweeps ran"}
    if is_stencil:
        # BASELINE's target (>= 70 % of the HBM roofline for the fused step at 512^3) is stated for the two-pass
        # step of SURVEY 8d, 48*N bytes: the same step time expressed in that accounting
        survey_gbs = 48.0 * N / ms_per_step / 1e6
        fused["survey_48N_accounting"] = {"bytes_per_step": 48.0 * N, "equivalent_gbs": survey_gbs if plain else None,
                                          "frac_of_measured_peak": survey_gbs / peak if plain else None,
                                          "frac_of_8TBs_nominal": survey_gbs / 8000.0 if plain else None}

    n_solves_e = Re * len(chunks)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "grid": list(grid), "unknowns": M_total,
                   "unknowns_per_gpu": M_local, "lanczos_m": chunks, "chunked": len(chunks) > 1, "reorth": wl["reorth"],
                   "cgs_passes": wl["cgs_passes"], "l2": "inputs larger than L2 (each vector %.2f GB)" % (8 * M_local / 1e9)
                   if 8 * M_local > 126e6 else "vectors of %.1f MB fit L2: this configuration is launch/latency-bound" % (8 * M_local / 1e6),
                   "sharding": ("z-slabs, one process per GPU" if is_stencil else "contiguous row blocks (z-slabs of cells), ghost-index exchange over peer memory, one process per GPU") if world > 1 else "single GPU",
                   "aggregate": ("strong scaling: value = K / time of the one global solve" if strong else
                                 "value = n_gpus * K / time: each GPU advances its 512^3 shard K steps; "
                                 "the global (n_gpus x larger) solve advances K steps"),
                   "repeats": R,
                   "timed_region": "the K-step solve (pre-step, K steps, alpha/beta back to the host) repeated %d times "
                                   "back to back = %.2f s; value and ms_per_step are the sustained figures over the whole "
                                   "region, burst = the fastest single solve" % (R, ms_dev / 1e3),
                   "global_steps_per_sec": R * K / (ms_dev / 1e3), "host": numa},
        "timed_region_s": ms_dev / 1e3,
        "burst": {"value": scale_n * K / (burst_ms / 1e3), "ms_per_step": burst_ms / K,
                  "note": "fastest single K-step solve of the region (device time of the loop alone)"},
        "e2e": {"value": scale_n * Re * K / e_wall, "unit": UNIT, "h2d_bytes_per_step": 8.0 * M_local * len(chunks) / K,
                "d2h_bytes_per_step": (sum(3 * (8 * (c + 2) + 512) + 32 for c in chunks)) / K,
                "wall_s": e_wall, "solves": n_solves_e,
                "note": "every solve: start vector H2D from pinned host memory (%.2f GB), K steps, alpha/beta D2H, "
                        "H_eff and the lowest Ritz values on the host" % (8.0 * M_local / 1e9)},
        "gpu_launches": launches,
        "gpu_launches_per_solve": launches / max(1, R * len(chunks)),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": per_kernel,
        "kernels_sustained": {k: {"launches": kern_hot[k][1], "avg_ms": kern_hot[k][0] / kern_hot[k][1],
                                  "achieved_gbs": (per_kernel[k]["alg_bytes"] / (kern_hot[k][0] / kern_hot[k][1]) / 1e6) if k in per_kernel else None}
                              for k in kern_hot if kern_hot[k][1]},
        "gram_schmidt_ms": gs_ms,
        "reorth_steps": reorths,
        "fused_step": fused,
        "parity_check": parity,
        "ritz_lowest": [float(x) for x in theta[:4]],
    }
    if world > 1:
        # bytes this rank stores into its peers' exchange buffers per plain step (NVLink peer stores issued by
        # the kernels themselves): two boundary planes of the new vector (structured grids) or the ghost
        # entries its neighbours gather from it (sparse), plus 2 x 8 B x (world - 1) for the alpha / beta sums
        if is_stencil:
            sent = 2.0 * 8.0 * float(np.prod(grid[:-1]))
        else:
            sent = 8.0 * float(len(solver.plan.send_lists(rank)[0]))
        sent += 2 * 8.0 * (world - 1)
        line["nvlink"] = {"sent_bytes_per_step_per_gpu": sent, "achieved_gbs": sent / ms_per_step / 1e6,
                          "peak_gbs_per_direction": 900.0, "frac": sent / ms_per_step / 1e6 / 900.0,
                          "note": "halo / ghost / scalar exchange is a fraction of a percent of the link; it rides inside "
                                  "the producing kernels (no NCCL call, no copy-engine transfer on the data path)"}
    if not is_stencil:
        line["config"]["value_free_spmv"] = bool(value_free)
        line["config"]["windowed_spmv_granules"] = int(windowed)
        ib = int(idx_bytes)
        line["config"]["spmv_bytes"] = (f"{ib}*nnz + 24*N: every off-diagonal entry of the operator is equal, so the kernel reads column "
                                        f"indices only{' (16-bit offsets into the staged window of x)' if windowed else ''} "
                                        "(SURVEY 8d's 12*nnz + 16*N would be %.3f GB per launch)" % ((12.0 * nnz_true + 16.0 * N) / 1e9)
                                        ) if value_free else "12*nnz + 16*N (SURVEY 8d)"
        line["config"]["nnz_per_gpu"] = int(nnz_true)
        line["config"]["sell_stored_over_true"] = float(nnz_stored) / max(1, nnz_true)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_reference_leg(args.workload)
            except Exception as e:      # the baseline is a report, it must not sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
