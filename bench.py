#!/usr/bin/env python
"""bench.py -- dGSEM Euler DOF-updates/s per stage (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells CELLS_PER_SIDE]

Workload (BASELINE.json configs[1]): rising thermal bubble, 3-D synthetic hex mesh of n^3 elements (default
100^3 = 1.0e6 elements, order 4, 1.25e8 LGL nodes), FP64, diffusion + buoyancy on, one reference time step =
one explicit stage (SURVEY finding 1).  One step = one pass of the hot path (sweep A, ghost update, sweep B,
ghost update) over the whole mesh.  DOF-updates/s = 5 * nodes * steps / time.

One JSON line on stdout (rank 0).  `value` = device-resident throughput timed with CUDA events on the launching
stream; `e2e` = the same step driven through the C ABI with HOST buffers (pinned upload of rho,U,T + step +
download inside the timed region); `roofline` = algorithmic bytes of the two sweeps / event time against the
measured HBM copy bandwidth; `cpu_baseline` = the reference's own CPU code (oracle/_ref/fast/euler, OpenMP on
all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ORDER = 4
NP = (ORDER + 1) ** 3


# ------------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_node(kernel_info: str, visc: bool = True) -> dict:
    """Bytes per LGL node and sweep, two accountings:

    * `A`, `B`, `total` -- SURVEY.md 8(d)'s ALGORITHMIC bytes for the data flow that is implemented: every datum once per sweep that needs
      it, neighbour traces as L2 hits (0 B), no padding.  Metrics on the fly (straight-edged meshes, the bench): sweep A reads rho,U,theta
      (5) and writes rho_new, grad U, grad theta (13); sweep B reads rho_old, rho_new, U, theta, the gradients, rho_ref, p_ref (20) and
      writes U, theta (4); + 1.6 doubles per node and step of element geometry (8 corner vertices, face ids): 150.4 + 198.4 = 348.8 B at
      order 4.  Stored metrics (curved meshes) = the reference's own layout, 534.4 B.  `roofline.frac` uses THESE bytes.
    * `impl_A`, `impl_B`, `impl_total` -- what the kernels really have to move through DRAM: the 128/125 element padding, p' and S = |U|+c
      stored by one sweep and read by the next, and the face traces sweep A publishes and sweep B consumes (written long before they are
      read, so they travel through DRAM on both sides).  Reported next to the ncu DRAM traffic; not the roofline numerator."""
    n1 = ORDER + 1
    pad = 128.0 / NP                                          # element stride padded from 125 to 128 doubles
    g = 12 if visc else 0                                     # gradU(9) + gradT(3)
    tri = kernel_info.startswith("v4") and "on the fly" in kernel_info
    geo = 8 * (25.0 / NP + 3.0 / n1)                          # corner vertices + face ids, per node and sweep
    if tri:
        A = 8 * (5 + 1 + g) + geo
        B = 8 * (2 + 4 + g + 2 + 4) + geo
    else:
        A = 8 * (5 + 1 + g + 10) + 8 * 4 * 3.0 / n1
        B = 8 * (2 + 4 + g + 2 + 10 + 4) + 8 * 4 * 3.0 / n1
    if kernel_info.startswith("v4"):
        metrics = 0 if tri else 10
        npf = n1 * n1
        tbs = ((7 * npf + 15) // 16) * 16                     # doubles per face-trace block (trace_bs)
        per_elem = 560 + 6 * tbs * 8                          # element record + six trace blocks
        iA = 8 * ((5 + 1 + 1 + metrics) + (2 + g)) * pad + per_elem / NP      # rho,U,T,p_ref,S [,Jinv,cV] -> rho_new,p',grads + traces
        iB = 8 * ((7 + 2 + g + metrics) + 5) * pad + per_elem / NP            # rho_o,rho_n,U,T,p',S,rho_ref,grads [,Jinv,cV] + traces -> U,T,S
        flow = "v4: " + ("metrics on the fly" if tri else "stored metrics") + ", dense face traces through DRAM, S = |U|+c kept with the state"
    else:
        face_tab = 6 * (4 + 4 + 24 + 24) / NP                 # faceOther, faceMeta, faceVec, faceUnit per element
        iA = 8 * ((5 + 10 + 1) + (2 + g)) * pad + face_tab
        iB = 8 * ((2 + 4 + 1 + g + 10 + 1) + 4) * pad + face_tab
        flow = "v1/v2: stored metrics, neighbour data counted as L2 hits"
    return dict(A=A, B=B, total=A + B, impl_A=iA, impl_B=iB, impl_total=iA + iB, flow=flow)


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons DURING the timed region (B200_PROFILING.md recipe), polled every few
    milliseconds from NVML -- the source `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` reads; a piped nvidia-smi
    block-buffers its output and loses most samples of a 0.2 s region.  Falls back to one nvidia-smi query per sample."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device = device
        self.samples = []            # (sm_mhz, sm_max_mhz, power_w, reasons:set)
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = int(vis.split(",")[device]) if vis and vis.split(",")[device].isdigit() else device
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml, self.handle = pynvml, h
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _one_nvml(self):
        n, h = self.nvml, self.handle
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        try:
            pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            pw = float("nan")
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        rs = set()
        for name, attr in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
            if bits & getattr(n, attr, 0):
                rs.add(name)
        return sm, self.smax, pw, rs

    def _one_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
        f = [x.strip() for x in out.split(",")]
        rs = {name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]) if val.lower().startswith("active")}
        return float(f[0]), float(f[1]), float(f[2]), rs

    def _loop(self):
        one = self._one_nvml if self.nvml else self._one_smi
        while not self.stop_flag.is_set():
            try:
                self.samples.append(one())
            except Exception:
                pass
            self.stop_flag.wait(0.004)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(x[0] for x in self.samples)
        reasons = set().union(*[x[3] for x in self.samples])
        pw = [x[2] for x in self.samples if x[2] == x[2]]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(x[1] for x in self.samples), "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi", "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference solver on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------------
def run_reference_sample(n_side: int, steps: int, warmup: int, threads: int | None = None) -> dict:
    from oracle import cases, run_ref          # test infrastructure, allowed here (cpu_baseline / reference arm only)
    variant = "fast" if run_ref.have_ref("fast") else "parity"
    if not run_ref.have_ref(variant):
        return run_oracle_port_sample(steps, warmup)
    threads = threads or os.cpu_count() or 1
    d = tempfile.mkdtemp(prefix="nsem_ref_")
    try:
        total = warmup + steps
        c = cases.bubble3d(n=n_side, order=ORDER, scheme="AB1")
        c.write(d, end_step=total, write_interval=10 * total + 7)       # no field dump inside the timed run
        wall, out = run_ref.run_euler(d, variant=variant, threads=threads, timeout=3000)
        stamps = [int(m.group(1)) for m in re.finditer(r"^(\d+) \[0\] Time ", out, flags=re.M)]
        end = re.search(r"^(\d+) \[0\] Exiting", out, flags=re.M)
        if len(stamps) < total or not end:
            return {"unavailable": "could not parse the reference log"}
        stamps.append(int(end.group(1)))
        t_ms = stamps[total] - stamps[warmup]
        nodes = n_side ** 3 * NP
        val = 5.0 * nodes * steps / (t_ms * 1e-3)
        return {"value": val, "unit": "DOF-updates/s", "cores": threads, "kind": "reference",
                "sample": f"bubble3d {n_side}^3 elements order {ORDER} ({nodes} nodes), {steps} steps after {warmup} warm-up, "
                          f"oracle/_ref/{variant}/euler (unmodified reference, 1 rank x {threads} OpenMP threads; no MPI in the image)",
                "ms_per_step": t_ms / steps}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_oracle_port_sample(steps: int, warmup: int, n_side: int = 4) -> dict:
    """No compiled reference on this box (oracle/_ref is built where /root/reference exists and travels with the repo): time the numpy
    restatement instead ("port", one thread, a much smaller sample -- it is a checker, not a fast CPU code)."""
    from oracle import case as ocase
    from oracle import cases
    d = tempfile.mkdtemp(prefix="nsem_port_")
    try:
        cases.bubble3d(n=n_side, order=ORDER, scheme="AB1").write(d, end_step=warmup + steps)
        orc = ocase.load_case(d, exact_order=False)
        orc.run(warmup)
        t0 = time.time()
        orc.run(steps)
        t_ms = (time.time() - t0) * 1e3
        nodes = n_side ** 3 * NP
        return {"value": 5.0 * nodes * steps / (t_ms * 1e-3), "unit": "DOF-updates/s", "cores": 1, "kind": "port",
                "sample": f"bubble3d {n_side}^3 elements order {ORDER} ({nodes} nodes), {steps} steps after {warmup} warm-up, numpy oracle "
                          f"(oracle/euler.py; oracle/_ref absent on this box)", "ms_per_step": t_ms / steps}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ------------------------------------------------------------------------------------------------------
def broadcast_unique_id(rank: int) -> bytes:
    """A fresh ncclUniqueId from rank 0 for one nsem context group (torch.distributed is only the plumbing that carries it)."""
    import ctypes

    import torch
    import torch.distributed as dist
    from nebulasem_b200 import capi
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        if capi.load_library().nsem_get_unique_id(buf) != 0:
            raise SystemExit("bench.py: ncclGetUniqueId failed")
        uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().tolist())


def mp_parity_check(rank: int, world: int, device: int, decomp: str, pgrid, nsteps: int = 10) -> dict:
    """Multi-GPU correctness where the driver can see it: the SAME small rising-bubble case stepped as `world` partitions (one per GPU,
    halo exchange as in the timed run) and as ONE partition on rank 0's GPU; the gathered fields must agree bit for bit (the face sums
    of an element have a fixed order and both sides of a cut face see identical traces).  Matches ASYNC_COMM,
    field.h:2255-2324 (with the gradients exchanged too, SURVEY finding 5)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from nebulasem_b200 import host
    n = (4 * pgrid[0], 4 * pgrid[1], 4 * pgrid[2])
    uid = broadcast_unique_id(rank)
    s = host.Solver.synthetic_part("bubble3d", n[0], n[1], n[2], ORDER, rank, world, decomp, pgrid)
    s.attach(device, rank, world, uid)
    s.step(nsteps)
    s.download()
    rho, U, T, _ = s.state()
    nb = s.gBCSfield
    cg = torch.tensor(s.u32("cellGlobal").astype(np.int64), device="cuda")
    ncell = n[0] * n[1] * n[2]
    out = torch.zeros((ncell, NP, 5), dtype=torch.float64, device="cuda")
    out[cg] = torch.tensor(np.concatenate([rho[:nb, None], U[:nb], T[:nb, None]], axis=1).reshape(s.nBCS, NP, 5), device="cuda")
    dist.all_reduce(out)                  # the partitions are disjoint: the sum assembles the global field
    kinfo = s.kernel_info
    s.close()
    res = None
    if rank == 0:
        ref = host.Solver.synthetic("bubble3d", n[0], n[1], n[2], ORDER)
        ref.attach(device)
        ref.step(nsteps)
        ref.download()
        r1, U1, T1, _ = ref.state()
        nb1 = ref.gBCSfield
        ref.close()
        one = np.concatenate([r1[:nb1, None], U1[:nb1], T1[:nb1, None]], axis=1).reshape(ncell, NP, 5)
        got = out.cpu().numpy()
        res = {"bitwise": bool(np.array_equal(got, one)), "max_abs_diff": float(np.abs(got - one).max()),
               "case": f"bubble3d {n[0]}x{n[1]}x{n[2]} elements order {ORDER}, {nsteps} steps, {world} {decomp} partitions vs 1 partition",
               "kernels": kinfo}
    flag = torch.tensor([1 if (res is None or res["bitwise"]) else 0], device="cuda")
    dist.broadcast(flag, 0)
    if int(flag.item()) == 0:
        raise SystemExit(f"bench.py: {world} partitions differ from one partition: {res}")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", dest="n", type=int, default=100, help="elements per side of the cubic mesh")
    ap.add_argument("--ref-cells", dest="ref_n", type=int, default=16, help="elements per side of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="bubble", choices=["bubble", "hill"],
                    help="bubble = BASELINE configs[1] (default, the metric's configuration); hill = configs[3], flow over a cosine hill on a "
                         "terrain-following mesh of the same element count (non-affine elements, DIRICHLET inlet)")
    ap.add_argument("--decomp", default="METIS", choices=["METIS", "XYZ", "CELLID"], help="domain decomposition for --gpus > 1")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's scaling run): --cells^3 elements PER GPU, the domain grows with the process grid; strong: ONE "
                         "mesh of about 4.0e6 elements (BASELINE configs[3]: bubble 160^3, hill 232x120x144) divided among the GPUs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(3, args.warmup)
    config = {"workload": f"rising thermal bubble 3-D synthetic hex mesh, order {ORDER}, {args.n}^3 = {args.n ** 3} elements per GPU "
                          f"({args.n ** 3 * NP} LGL nodes), diffusion+buoyancy on, one forward-Euler stage per step",
              "elements_per_gpu": args.n ** 3, "order": ORDER, "l2": "state and metrics (tens of GB) far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_sample(args.ref_n, steps, warmup)
        # the reference cannot hold the 100^3 mesh (15.9 GB per MeshMatrix object): its arm runs a bounded sample of the
        # same workload and says so where the driver compares configurations
        config = dict(config)
        config["workload"] = (f"rising thermal bubble 3-D synthetic hex mesh, order {ORDER}, SAMPLE of {args.ref_n}^3 = {args.ref_n ** 3} elements "
                              f"({args.ref_n ** 3 * NP} LGL nodes) of the {args.n}^3-element workload, diffusion+buoyancy on, one forward-Euler "
                              f"stage per step; per-DOF throughput on the host cores")
        config["elements_per_gpu"] = args.ref_n ** 3
        config["sample_of"] = f"{args.n}^3 elements per GPU"
        line = {"impl": "reference", "metric": "dGSEM Euler DOF-updates/s per stage", "unit": "DOF-updates/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config}
        if "unavailable" in r:
            line["unavailable"] = r["unavailable"]
        else:
            line.update({"value": r["value"], "ms_per_step": r["ms_per_step"], "cpu_baseline": r,
                         "e2e": {"value": r["value"], "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line), flush=True)
        return 0

    # torchrun exports OMP_NUM_THREADS=1; the host-side mesh/geometry set-up (not timed) is OpenMP-parallel
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, world)))
    import numpy as np
    import torch
    from nebulasem_b200 import host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    device = local_rank
    torch.cuda.set_device(device)
    import torch.distributed as dist
    pgrid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (world, 1, 1))
    uid_bytes = None
    mp_parity = None
    if world > 1:
        # one process per GPU; torch.distributed is plumbing (rendezvous, barrier, max-over-ranks), the halo is NCCL
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        mp_parity = mp_parity_check(rank, world, device, args.decomp, pgrid)
        uid_bytes = broadcast_unique_id(rank)
        config["parallelism"] = (f"{world} partitions ({args.decomp}), one per GPU, global mesh {args.n * pgrid[0]}x{args.n * pgrid[1]}x"
                                 f"{args.n * pgrid[2]} elements on a {1000 * pgrid[0]}x{1000 * pgrid[1]}x{1000 * pgrid[2]} m domain (element size as at N=1), face-trace halo")

    t0 = time.time()
    # The example's dt = 0.00125 belongs to its 6^3 (+2 AMR levels, ~42 m) mesh; one step is ONE forward-Euler stage, so the
    # finer benchmark mesh keeps the example's acoustic Courant number instead of its dt (at dt = 0.00125 the 100^3 mesh
    # diverges after ~30 steps, in the reference's scheme as in this one).  The throughput does not depend on dt.
    dt = min(0.00125, 0.00125 * 24.0 / args.n)
    strong = args.scaling == "strong"
    dom = pgrid                                                            # bubble domain in km: grows with the process grid (weak)
    gn = (args.n * pgrid[0], args.n * pgrid[1], args.n * pgrid[2])         # global bubble mesh: n^3 elements per GPU (weak)
    if strong:
        # ONE mesh of ~4.0e6 elements whatever the GPU count (--cells 100 -> 160^3 on the 1 km cube; smaller --cells for trials)
        side = max(8, round(160 * args.n / 100.0 / 8) * 8)
        gn, dom = (side, side, side), (1, 1, 1)
        dt = min(0.00125, 0.00125 * 24.0 / side)
    if args.workload == "hill":
        # BASELINE configs[3]: the hills/hill block layout extruded to 3-D, 232 x 30 x 144 elements per GPU at --cells 100,
        # extruded further in y for more GPUs (weak) or 232 x 120 x 144 = 4.0e6 elements in all (strong); dt keeps the Courant number of
        # the parity case (nz = 8, dt = 0.001)
        f = args.n / 100.0
        ny = max(2, round(120 * f)) if strong else max(2, round(30 * f)) * world
        gn = (max(4, round(232 * f)), max(world, ny // world * world), max(4, round(144 * f)))
        dt = min(0.001, 0.001 * 8.0 / gn[2])
    total = gn[0] * gn[1] * gn[2]
    if args.workload == "hill":
        config["workload"] = (f"flow over a cosine hill, terrain-following 3-D hex mesh {gn[0]}x{gn[1]}x{gn[2]} = {total} elements "
                              f"({total // world} per GPU), order {ORDER}, U = (10,0,0) DIRICHLET inlet, diffusion+buoyancy on, "
                              f"one forward-Euler stage per step")
    elif strong:
        config["workload"] = (f"rising thermal bubble 3-D synthetic hex mesh, order {ORDER}, {gn[0]}^3 = {total} elements in all "
                              f"({total // world} per GPU, {total * NP} LGL nodes), diffusion+buoyancy on, one forward-Euler stage per step")
    if args.workload == "hill" or strong:
        config["elements_per_gpu"] = total // world
        if world > 1:
            config["parallelism"] = f"{world} partitions ({args.decomp}), one per GPU, of the {gn[0]}x{gn[1]}x{gn[2]} mesh, face-trace halo"
    config["dt"] = dt
    config["scaling"] = args.scaling
    kind = f"hill3d:{dt!r}" if args.workload == "hill" else f"bubble3d:{dom[0]},{dom[1]},{dom[2]},{dt!r}"
    part = (1, world, 1) if args.workload == "hill" else pgrid
    if world == 1:
        s = host.Solver.synthetic(kind, gn[0], gn[1], gn[2], ORDER)
        s.attach(device)
    else:
        s = host.Solver.synthetic_part(kind, gn[0], gn[1], gn[2], ORDER, rank, world, args.decomp, part)
        s.attach(device, rank, world, uid_bytes)
    t_setup = time.time() - t0
    nodes_local = s.gBCSfield
    launches0 = s.launch_count

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s.step(warmup)
    s.sync()
    barrier()
    sampler = ClockSampler(device)
    sampler.start()
    if world > 1:
        s.halo_wait_ms()                                     # reset the wait counters
    ms, _ = s.time_steps(steps, per_kernel=False)
    halo_wait = s.halo_wait_ms() if world > 1 else None      # this rank's time inside the timed region spent waiting for its neighbours
    barrier()
    clocks = sampler.stop()
    clocks_per_rank = None
    if world > 1:
        # every rank samples ITS GPU during the timed region: the job is paced by the slowest partition twice per step, so one GPU that the
        # board's power management holds below the others shows up as lost scaling efficiency
        smp = sorted(x[0] for x in sampler.samples) or [0.0]
        pw = [x[2] for x in sampler.samples if x[2] == x[2]] or [0.0]
        capped = sum(1 for x in sampler.samples if "sw_power_cap" in x[3])
        mine = torch.tensor([smp[len(smp) // 2], smp[0], max(pw), capped / max(1, len(sampler.samples)), halo_wait[0] / steps, halo_wait[1] / steps],
                            dtype=torch.float64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        clocks_per_rank = {"columns": ["sm_mhz_median", "sm_mhz_min", "power_w_max", "fraction_of_samples_power_capped",
                                       "wait_for_neighbours_after_A_ms_per_step", "wait_for_neighbours_after_B_ms_per_step"],
                           "rows": [[round(float(v), 3) for v in t.tolist()] for t in allc]}
    launches = s.launch_count - launches0
    if world > 1:
        config["halo"] = s.halo_info
    nodes = nodes_local
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # device time, max over ranks
        ms = float(t.item())
        nn = torch.tensor([nodes_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(nn)
        nodes = int(nn.item())
    steps_per_launchcount = launches / max(1, warmup + steps)
    launches = int(round(steps_per_launchcount * steps))   # kernels inside the timed region
    pk_steps = max(2, steps // 2)
    ms_pk_total, pk = s.time_steps(pk_steps, per_kernel=True)
    barrier()
    per_rank = None
    if world > 1:
        # every rank's own kernel times (sweep A, ghost update + halo, sweep B, ghost update + halo; ms per step) and element count: the step
        # is paced by the slowest partition twice per step (one halo after each sweep)
        mine = torch.tensor([pk[0] / pk_steps, pk[1] / pk_steps, pk[2] / pk_steps, pk[3] / pk_steps, float(s.nBCS), float(s.nCells - s.nBCS)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 4) for v in t.tolist()] for t in allr]
    value = 5.0 * nodes * steps / (ms * 1e-3)

    # finite-state sanity after all those steps (a diverged run would be a meaningless number)
    s.download()
    rho, U, T, p = s.state()
    if not (np.isfinite(rho).all() and np.isfinite(U).all() and np.isfinite(T).all()) and not os.environ.get("NSEM_EXPERIMENT_ALLOW_NONFINITE"):
        raise SystemExit("bench.py: state is not finite after the timed steps")

    # roofline of the dominant kernel (the sweep with the larger share of the step) and of the pair
    kinfo = s.kernel_info
    bpn = algorithmic_bytes_per_node(kinfo, visc=True)
    peak, peak_src = measured_peak_gbs()
    tA, tB = pk[0] / pk_steps * 1e-3, pk[2] / pk_steps * 1e-3
    achieved_B = bpn["B"] * nodes_local / tB / 1e9
    achieved_A = bpn["A"] * nodes_local / tA / 1e9
    dom = "A" if tA >= tB else "B"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("cells") == args.n and tj.get("kernels", "") == kinfo.split()[0]:
            traffic = tj[f"sweep{dom}_bytes_per_launch"]   # ncu dram__bytes_read+write of the dominant kernel, per launch
    kname = {"A": "v4::sweepA_v4<5,5,5,visc,tri>", "B": "v4::sweepB_v4<5,5,5,visc,tri>"}[dom] if kinfo.startswith("v4") else f"sweep{dom} ({kinfo})"
    step_s = ms / steps * 1e-3
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved_A if dom == "A" else achieved_B, "peak": peak, "unit": "GB/s",
                "frac": (achieved_A if dom == "A" else achieved_B) / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bpn[dom] * nodes_local, "peak_source": peak_src, "kernels": kinfo, "data_flow": bpn["flow"],
                "bytes_per_node": {"sweepA": bpn["A"], "sweepB": bpn["B"], "step": bpn["total"],
                                   "what": "SURVEY 8(d) algorithmic bytes of the implemented flow (neighbour traces = L2 hits, no padding)"},
                "sweepA": {"achieved": achieved_A, "frac": achieved_A / peak, "ms": tA * 1e3},
                "sweepB": {"achieved": achieved_B, "frac": achieved_B / peak, "ms": tB * 1e3},
                "bc_ms": [pk[1] / pk_steps, pk[3] / pk_steps],
                "step": {"achieved": bpn["total"] * nodes / world / step_s / 1e9, "frac": bpn["total"] * nodes / world / step_s / 1e9 / peak,
                         "note": "per GPU"},
                # second accounting: what the kernels must really move through DRAM (padding, p', S, face traces written and re-read)
                "implemented": {"bytes_per_node": {"sweepA": bpn["impl_A"], "sweepB": bpn["impl_B"], "step": bpn["impl_total"]},
                                "sweepA_frac": bpn["impl_A"] * nodes_local / tA / 1e9 / peak, "sweepB_frac": bpn["impl_B"] * nodes_local / tB / 1e9 / peak,
                                "step_frac": bpn["impl_total"] * nodes / world / step_s / 1e9 / peak}}

    # e2e: host buffers in, host buffers out, every step -- through the pipelined entry points of the C ABI (nsem_upload_state_async /
    # nsem_euler_step / nsem_download_state_async): every step's input batch is copied from pinned host memory and its result copied back
    # to pinned host memory inside the timed region; the download of step k overlaps the upload of step k+1 (PCIe is full duplex).
    # The strictly serial variant (upload, step, download, each blocking) is reported next to it.
    e2e = None
    if not args.no_e2e:
        e_steps = max(1, min(steps, 3))
        gall = s.gALL

        def wall_max(x):
            # every rank moves its own partition through the C ABI with its own host buffers: the job's time is the slowest rank's
            if world == 1:
                return x
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        s.upload(); s.step(1); s.download()                # warm the transfer path (first call page-locks the host arrays)
        barrier()
        t1 = time.perf_counter()
        for _ in range(e_steps):
            s.upload()
            s.step(1)
            s.download()
        torch.cuda.synchronize()
        dt_serial = wall_max(time.perf_counter() - t1)
        s.upload_async(); s.step(1); s.download_async(); s.sync()      # allocates the staging buffers and page-locks the output arrays
        p_steps = 2 * e_steps
        barrier()
        t1 = time.perf_counter()
        for _ in range(p_steps):
            s.upload_async()
            s.step(1)
            s.download_async()
        s.sync()
        dt_e = wall_max(time.perf_counter() - t1)
        rho_o, U_o, T_o, p_o = s.state_out()
        if not (np.isfinite(rho_o).all() and np.isfinite(U_o).all() and np.isfinite(T_o).all()):
            raise SystemExit("bench.py: pipelined e2e returned a non-finite state")
        e2e = {"value": 5.0 * nodes * p_steps / dt_e, "unit": "DOF-updates/s", "h2d_bytes_per_step": 6 * gall * 8 * world,
               "d2h_bytes_per_step": 6 * gall * 8 * world, "steps": p_steps, "ms_per_step": dt_e / p_steps * 1e3,
               "what": "nsem_upload_state_async(rho,U,T,p pinned host arrays) + nsem_euler_step(1) + nsem_download_state_async per step, "
                       "nsem_sync at the end, wall clock; copies of consecutive steps overlap",
               "serial": {"value": 5.0 * nodes * e_steps / dt_serial, "ms_per_step": dt_serial / e_steps * 1e3, "steps": e_steps,
                          "what": "nsem_upload_state + nsem_euler_step(1) + nsem_download_state, each blocking"}}

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cpu = run_reference_sample(args.ref_n, 5, 1)
    if world > 1:
        dist.barrier()
    if rank != 0:
        s.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {"metric": "dGSEM Euler DOF-updates/s per stage", "value": value, "unit": "DOF-updates/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "node_updates_per_s": value / 5.0, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "setup_s": t_setup}
    if mp_parity is not None:
        line["mp_parity"] = mp_parity
    if clocks_per_rank is not None:
        line["clocks_per_rank"] = clocks_per_rank
    if per_rank is not None:
        line["per_rank"] = {"columns": ["sweepA_ms", "bcA_halo_ms", "sweepB_ms", "bcB_halo_ms", "elements", "ghost_cells"], "rows": per_rank}
    print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
