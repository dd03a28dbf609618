#!/usr/bin/env python
"""bench.py -- RDF pair-distances/s and MSD / Green-Kubo atom-lag updates/s on B200.

Workload (BASELINE.json configs[4], the scaling-sweep system): a 1,000,000-atom two-species
melt (2 x 500,000 atoms, rho = 0.05 / A^3 => L = 271.44 A, default cutoff L/2 - 0.1 and
int(cutoff / 0.01) = 13,562 bins), 2,000 frames, data_range = 500.  One *step* on one GPU is

  RDF phase       one sampled frame of the full system: pack -> mdk_rdf_hist
                  (4.999995e11 pair distances);  frames shard across ranks
  dynamics phase  a 125,000-atom shard (1/8 of the system; atoms shard across ranks) over all
                  2,000 frames: unwrap -> MSD (W = 1500 windows x 500 lags) and
                  velocity ACF (same update count) -> per-window series for the SEM

followed, for N > 1, by the NCCL all-reduce of the histograms and series.  Scaling is weak:
per-GPU work is fixed (at N = 8 the dynamics phase covers exactly the 1,000,000 atoms).

The ONE JSON line carries the primary metric (RDF pair-distances/s) plus ``secondary``
entries for the MSD, ACF, unwrap and ionic-current kernels, each with its own roofline.
``--impl reference`` times the CPU restatement of the reference algorithm (oracle/) on the
host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "rdf_pair_distances_per_s"
UNIT = "pair-distances/s"
N_ATOMS = 1_000_000
N_SPECIES_ATOMS = 500_000
DENSITY = 0.05
N_FRAMES = 2000
DATA_RANGE = 500
DYN_SHARD = 125_000
BOX = (N_ATOMS / DENSITY) ** (1.0 / 3.0)
FLOP_PER_PAIR = 20.0
FLOP_PER_MSD = 9.0
FLOP_PER_ACF = 6.0


def workload_config(n_gpus, small=False):
    return {
        "workload": "C5: 1,000,000-atom two-species melt (2x500k), 2,000 frames, "
                    "RDF (13,562 bins, 3 species pairs) + Einstein MSD / GK ACF data_range=500",
        "rdf_frames_per_gpu_per_step": 1,
        "dynamics_atoms_per_gpu_per_step": DYN_SHARD,
        "dynamics_frames": N_FRAMES,
        "data_range": DATA_RANGE,
        "correlation_time": 1,
        "box": BOX,
        "sharding": f"rdf by frame, dynamics by atom, {n_gpus} rank(s)",
        "l2": "flushed between timed steps (256 MiB write); dynamics inputs (3 GB) exceed L2",
        "reduced_size_debug_run": bool(small),
    }


# ----------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ----------------------------------------------------------------------------------------
# CPU baseline: the oracle (reference algorithm restated) on the host cores
# ----------------------------------------------------------------------------------------
def _cpu_rdf_task(args):
    seed, n_atoms = args
    from oracle import rdf as orc

    rng = np.random.default_rng(seed)
    L = (n_atoms / DENSITY) ** (1.0 / 3.0)
    half = n_atoms // 2
    pos = {"A": (rng.random((half, 1, 3)) * L).astype(np.float32),
           "B": (rng.random((half, 1, 3)) * L).astype(np.float32)}
    box = np.array([L, L, L])
    cutoff = orc.default_cutoff(box)
    nbins = orc.default_number_of_bins(cutoff)
    t0 = time.perf_counter()
    # reference plan: one frame per batch, atom minibatches of 100
    orc.rdf_counts(pos, ["A", "B"], box, np.arange(1), cutoff, nbins, 100, 1)
    return time.perf_counter() - t0, n_atoms * (n_atoms - 1) // 2


def _cpu_dyn_task(args):
    seed, kind, n_atoms, n_frames, data_range = args
    from oracle import dynamics as od

    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(0, 0.05, size=(n_atoms, n_frames, 3)), axis=1).astype(np.float32)
    plan = dict(batch_size=n_frames, n_batches=1, remainder=0, minibatch=False)
    t0 = time.perf_counter()
    if kind == "msd":
        od.einstein_msd(x, plan, data_range, 1, np.arange(data_range))
    else:
        od.gk_diffusion_acf(x, plan, data_range, 1, np.arange(data_range) * 1.0, 1.0, 1.0)
    return time.perf_counter() - t0, (n_frames - data_range) * n_atoms * data_range


def cpu_reference(workers: int, rdf_atoms=8000, msd_atoms=160, acf_atoms=32):
    """Times the oracle on `workers` processes (each runs an independent sample of the same
    workload shape).  Returns dict metric -> (units/s, sample description)."""
    import multiprocessing
    from concurrent.futures import ProcessPoolExecutor

    out = {}
    # spawn, not fork: the parent may hold a CUDA context, and a forked child that garbage
    # collects inherited CUDA tensors dies with an initialisation error
    with ProcessPoolExecutor(max_workers=workers,
                             mp_context=multiprocessing.get_context("spawn")) as ex:
        def run(task, argl):
            t0 = time.perf_counter()
            res = list(ex.map(task, argl))
            wall = time.perf_counter() - t0
            return sum(u for _, u in res) / wall

        out["rdf"] = (run(_cpu_rdf_task, [(s, rdf_atoms) for s in range(workers)]),
                      f"{workers} x 1 frame of a {rdf_atoms}-atom two-species system at the same "
                      "density, reference plan (1-frame batches, 100-atom minibatches, 3 masked "
                      "species-pair passes)")
        out["msd"] = (run(_cpu_dyn_task, [(s, "msd", msd_atoms, N_FRAMES, DATA_RANGE)
                                          for s in range(workers)]),
                      f"{workers} x {msd_atoms} atoms x {N_FRAMES} frames, data_range {DATA_RANGE}, "
                      "per-window loop")
        out["acf"] = (run(_cpu_dyn_task, [(s, "acf", acf_atoms, N_FRAMES, DATA_RANGE)
                                          for s in range(workers)]),
                      f"{workers} x {acf_atoms} atoms x {N_FRAMES} frames, data_range {DATA_RANGE}, "
                      "per-window complex128 FFT autocorrelation")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    scale = 0.25 if args.small else 1.0
    vals, t_all = [], []
    res = None
    for it in range(warm + steps):
        t0 = time.perf_counter()
        res = cpu_reference(cores, rdf_atoms=int(4000 * scale) or 500, msd_atoms=int(48 * scale) or 8,
                            acf_atoms=int(8 * scale) or 2)
        if it >= warm:
            vals.append(res)
            t_all.append(time.perf_counter() - t0)
    rdf = statistics.mean(v["rdf"][0] for v in vals)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rdf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * statistics.mean(t_all), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.small),
        "cpu_baseline": {"value": rdf, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": res["rdf"][1]},
        "secondary": [
            {"metric": "msd_atom_lag_updates_per_s",
             "value": statistics.mean(v["msd"][0] for v in vals), "unit": "atom-lag updates/s",
             "sample": res["msd"][1]},
            {"metric": "acf_atom_lag_updates_per_s",
             "value": statistics.mean(v["acf"][0] for v in vals), "unit": "atom-lag updates/s",
             "sample": res["acf"][1]},
        ],
        "e2e": {"value": rdf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference TF path is not installable here (no tensorflow / tfp / h5py wheels); "
                "this is the line-faithful NumPy restatement in oracle/ on all host cores",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import RdfEngine, acf_series, msd_series, plan_windows
    from lammps_analysis_b200.synthetic import device_fluid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # NCCL's INFO log is the evidence of which ranks joined the communicator: keep whatever the
        # launcher configured, otherwise send it to stderr (stdout stays the one JSON line)
        # (the GPU image presets NCCL_DEBUG to a quieter level, which would hide the rank lines)
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    if args.strong_only:
        def _barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        rec = run_strong_leg(args, world, rank, dev, _barrier)
        if rank == 0:
            print(json.dumps({"strong": rec}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    steps, warm = max(1, args.steps), max(3, args.warmup)
    n_iter = steps + warm
    small = args.small
    n_sp = N_SPECIES_ATOMS if not small else 20_000
    n_frames = N_FRAMES if not small else 600
    shard = DYN_SHARD if not small else 4_000
    N = DATA_RANGE if not small else 100
    box_l = BOX if not small else (2 * n_sp / DENSITY) ** (1 / 3)
    box = np.array([box_l] * 3)
    cutoff = box_l / 2 - 0.1
    nbins = int(cutoff / 0.01)

    # ---- synthetic inputs, resident in HBM -------------------------------------------------
    rdf_frames_total = n_iter
    sp_traj = [device_fluid(n_sp, rdf_frames_total, box_l, 500 + 10 * rank + s, dev)
               for s in range(2)]
    pos = device_fluid(shard, n_frames, box_l, 700 + rank, dev, sigma_step=0.4)
    gen = torch.Generator(device=dev)
    gen.manual_seed(900 + rank)
    vel = torch.randn(shard, n_frames, 3, device=dev, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    unw = torch.empty_like(pos)
    carry_img = torch.zeros(shard, 3, dtype=torch.float64, device=dev)
    plan = dict(batch_size=n_frames, n_batches=1, remainder=0, minibatch=False)
    launches = plan_windows(plan, N, 1, shard)
    W = launches[0][4]
    tau = np.arange(N)
    J = torch.zeros(n_frames, 3, dtype=torch.float64, device=dev)

    eng = RdfEngine([n_sp, n_sp], box, cutoff, nbins, drop_first=True, device=dev)
    pairs_per_frame = eng.pairs_per_frame()
    upd_per_step = W * shard * N

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(i, rec=None):
        """One pass of the hot path, inputs resident in HBM.  rec: dict of event lists."""
        marks = {}

        def mark(name):
            if rec is not None:
                e = ev()
                e.record()
                marks[name] = e

        mark("t0")
        eng.hist.zero_()
        eng.record_events = rec is not None
        # Hilbert-ordered pack (CUB radix sort + gather), tile boxes, then the pair kernel
        eng.add_frames(sp_traj, [i], check_extent=True)   # extent check enables the wrapped path
        eng.record_events = False
        if world > 1:
            dist.all_reduce(eng.hist)
        mark("rdf_end")
        carry_img.zero_()
        mark("k_unw0")
        K.unwrap(pos, box, None, carry_img, False, unw)
        mark("k_unw1")
        msd, _ = msd_series(unw, launches, N, 1, tau)
        mark("k_msd1")
        acf, _, wins, _ = acf_series(vel, launches, N, 1, per_window=True)
        mark("k_acf1")
        J.zero_()
        mark("k_ion0")
        K.ionic_current(vel, 1.0, J)
        mark("k_ion1")
        if world > 1:
            dist.all_reduce(msd)
            dist.all_reduce(acf)
            dist.all_reduce(J)
        mark("t1")
        if rec is not None:
            rec.append(marks)
        return msd, acf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        step(i)
        flush.fill_(i & 255)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches_before = K.launch_count
    eng.kernel_events = []
    rec = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(warm, warm + steps):
        step(i, rec)
        flush.fill_(i & 255)      # L2 flush, outside the per-step event pairs
    barrier()
    wall = time.perf_counter() - wall0
    gpu_launches = K.launch_count - launches_before
    clocks = sampler.stop() if rank == 0 else None

    def span(a, b):
        return sum(m[a].elapsed_time(m[b]) for m in rec) * 1e-3  # seconds over all steps

    t = {
        "step": span("t0", "t1"), "rdf_phase": span("t0", "rdf_end"),
        "dyn_phase": span("rdf_end", "t1"),
        "rdf_kernel": sum(a.elapsed_time(b) for a, b in eng.kernel_events) * 1e-3,
        "unwrap_kernel": span("k_unw0", "k_unw1"), "msd_kernel": span("k_unw1", "k_msd1"),
        "acf_kernels": span("k_msd1", "k_acf1"), "ionic_kernel": span("k_ion0", "k_ion1"),
    }
    keys = sorted(t)
    tt = torch.tensor([t[k] for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = dict(zip(keys, tt.cpu().tolist()))

    # ---- end-to-end: the same step through the public calculator API, HOST-resident store ----
    # experiment "rdf": the two-species system, n_iter frames on the host; experiment "dyn": this
    # rank's 125,000-atom shard with wrapped Positions + Velocities on the host.  Every step
    # starts with nothing on the device: host -> device copies, the kernels, and the device ->
    # host reads of the results (and of the unwrapped positions the transformation persists)
    # are all inside the timed region.  The scipy post-processing (line fits, trapezoids) that
    # follows the hot path is outside it.
    import tempfile

    from lammps_analysis_b200 import distributed as mdk_dist
    from lammps_analysis_b200.config import config as mdk_config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project

    mdk_config.planner_memory_bytes = 60e9      # the SURVEY.md A.5 plan (C5: 2 atom batches)
    project = Project(f"bench{rank}", storage_path=tempfile.mkdtemp(prefix="mdk_bench_"),
                      persist=False, sharded=False)   # weak scaling: a private replica per rank
    exp_rdf = project.add_experiment("rdf", timestep=0.002, temperature=300.0, units="real")
    exp_rdf.add_data(ScriptInput({"A": {"Positions": sp_traj[0].cpu().numpy()},
                                  "B": {"Positions": sp_traj[1].cpu().numpy()}},
                                 box, atom_major=True))
    exp_dyn = project.add_experiment("dyn", timestep=0.002, temperature=300.0, units="real")
    exp_dyn.add_data(ScriptInput({"A": {"Positions": pos.cpu().numpy(),
                                        "Velocities": vel.cpu().numpy()}}, box, atom_major=True))
    exp_dyn.species["A"].charge = 1.0
    h2d_rdf = 2 * n_sp * 12
    d2h_rdf = eng.hist.numel() * 8
    h2d_dyn = 2 * shard * n_frames * 12
    d2h_dyn = shard * n_frames * 12 + (2 * N + W * N) * 8 + n_frames * 3 * 4

    def e2e_step(i, rec2):
        e0, e1, e2 = ev(), ev(), ev()
        exp_rdf.store.invalidate()
        exp_dyn.store.invalidate()
        for path in ("A/Unwrapped_Positions", "Observables/Ionic_Current"):
            if exp_dyn.store.check_existence(path):
                exp_dyn.store.remove(path)
        exp_rdf.version += 1          # defeat the result cache: every step recomputes
        exp_dyn.version += 1
        torch.cuda.synchronize()
        e0.record()
        # every rank owns its shard here (weak scaling): calculators run rank-local and the
        # exchange step is issued below
        with mdk_dist.local_only():
            rdf = exp_rdf.run.RadialDistributionFunction(start=i, stop=i,
                                                         number_of_configurations=1, plot=False)
        if world > 1:
            y = torch.tensor(np.array([rdf[k]["y"][1:] for k in rdf.keys()]), device=dev)
            dist.all_reduce(y)
        e1.record()
        with mdk_dist.local_only():
            msd_sum, acf_sum = e2e_dynamics()
        if world > 1:
            red = torch.tensor(np.stack([msd_sum, acf_sum]), device=dev)
            dist.all_reduce(red)
        e2.record()
        torch.cuda.synchronize()
        rec2.append((e0, e1, e2))
        return rdf, msd_sum, acf_sum

    def e2e_dynamics():
        # the public calls: Einstein resolves its Unwrapped_Positions dependency by running the
        # CoordinateUnwrapper transformation; line fit and trapezoid post-processing included
        ein = exp_dyn.run.EinsteinDiffusionCoefficients(data_range=N, plot=False)
        gk = exp_dyn.run.GreenKuboDiffusionCoefficients(data_range=N, plot=False)
        exp_dyn.run.IonicCurrent()
        exp_dyn.store.flush()   # the unwrapped positions' write-back (side stream) has landed
        return np.array(ein["A"]["msd"]), np.array(gk["A"]["acf"])

    e2e_step(0, [])
    barrier()
    rec2 = []
    for k in range(steps):
        e2e_step(1 + k % (n_iter - 1), rec2)
    barrier()
    e_rdf = sum(a.elapsed_time(b) for a, b, _ in rec2) * 1e-3
    e_dyn = sum(b.elapsed_time(c) for _, b, c in rec2) * 1e-3
    te = torch.tensor([e_rdf, e_dyn], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_rdf, e_dyn = te.cpu().tolist()

    # ---- HBM-bound regime of the correlation kernels (short lag ranges), same shard ---------------
    # not part of the step: data_range 500 is FP32-bound; these probes show the streaming kernels
    # against the HBM roofline north_star names for the correlation path
    def probe(fn, reps=5):
        for _ in range(3):
            fn()
            flush.fill_(3)
        tot = 0.0
        for _ in range(reps):
            e0, e1 = ev(), ev()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1) * 1e-3
            flush.fill_(4)
        return tot / reps

    probes = []
    for n_short in (2, 4, 8, 16, 32, 64, 100):
        if n_short >= N:
            continue
        l_short = plan_windows(plan, n_short, 1, shard)
        t_msd = probe(lambda: msd_series(unw, l_short, n_short, 1, np.arange(n_short)), reps=3)
        t_acf = probe(lambda: acf_series(vel, l_short, n_short, 1, per_window=False), reps=3)
        probes.append((n_short, t_msd, t_acf, l_short[0][4] * shard * n_short))

    # ---- strong scaling: the fixed C5 problem through the calculators' own sharding -------------
    spatial_sort = eng.spatial_sort
    del sp_traj, pos, vel, unw, eng, project, exp_rdf, exp_dyn
    torch.cuda.empty_cache()
    strong = None
    if not args.no_strong:
        strong = run_strong_leg(args, world, rank, dev, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline denominators ---------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback"
    fp32_peak = K.peak_fp32(True)
    fp32_src = "libmdk FFMA2 micro-benchmark, measured live (148 SM x 128 lanes x 2 x f_SM)"

    total_pairs = pairs_per_frame * steps * world
    total_upd = upd_per_step * steps * world
    total_af = shard * n_frames * steps * world
    value = total_pairs / t["rdf_phase"]
    rdf_tflops = FLOP_PER_PAIR * pairs_per_frame * steps / t["rdf_kernel"] * 1e-12

    def hbm_roof(bytes_per_launch, seconds_all_steps, traffic=None):
        ach = bytes_per_launch * steps / seconds_all_steps * 1e-9
        return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": hbm_src}

    def fp32_roof(flops_per_launch, seconds_all_steps, traffic=None):
        ach = flops_per_launch * steps / seconds_all_steps * 1e-12
        return {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak, "traffic": traffic, "peak_source": fp32_src}

    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(1, rdf_atoms=8000 if not small else 1000,
                            msd_atoms=160 if not small else 16,
                            acf_atoms=32 if not small else 4)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * t["step"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, small),
        "phases_ms_per_step": {k: 1e3 * v / steps for k, v in t.items()},
        "wall_s_timed_region": wall,
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload
        # (the bench's own 2 x 500k frame), ncu --set full capture profiles/r02ad_ncu_rdf_final.md
        "roofline": dict(fp32_roof(FLOP_PER_PAIR * pairs_per_frame, t["rdf_kernel"],
                                   traffic=None if small else 65.9e6),
                         traffic_unit="bytes per launch (ncu capture r02ad: 31.4 MB read + 34.5 MB "
                                      "written; 12 MB of coordinates are algorithmic, the rest is "
                                      "the flush of the per-CTA histograms -- the kernel is bound "
                                      "on chip, by the shared-memory pipe)",
                         kernel="rdf_pair_hist_kernel",
                         algorithmic="20 FLOP per pair-distance x 4.999995e11 pairs per launch "
                                     "(all i<j pairs count, including the blocks the kernel "
                                     "proves to lie beyond the cutoff and skips)",
                         spatial_sort=spatial_sort,
                         tflops=rdf_tflops),
        "e2e": {"value": pairs_per_frame * steps * world / e_rdf, "unit": UNIT,
                "h2d_bytes_per_step": h2d_rdf, "d2h_bytes_per_step": d2h_rdf,
                "api": "experiment.run.RadialDistributionFunction(start=f, stop=f, "
                       "number_of_configurations=1) on a host-resident store"},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        "secondary": [
            {"metric": "msd_atom_lag_updates_per_s", "unit": "atom-lag updates/s",
             "value": total_upd / t["msd_kernel"],
             "roofline": dict(fp32_roof(FLOP_PER_MSD * upd_per_step, t["msd_kernel"]),
                              kernel="msd_dense_kernel",
                              hbm=hbm_roof(12.0 * shard * n_frames, t["msd_kernel"])),
             "e2e": {"value": 2 * total_upd / e_dyn, "unit": "atom-lag updates/s (MSD+ACF)",
                     "h2d_bytes_per_step": h2d_dyn, "d2h_bytes_per_step": d2h_dyn,
                     "api": "run.EinsteinDiffusionCoefficients(data_range=500) (runs "
                            "CoordinateUnwrapper, fit included) + "
                            "run.GreenKuboDiffusionCoefficients(data_range=500) + "
                            "run.IonicCurrent() on a host-resident store"}},
            {"metric": "acf_atom_lag_updates_per_s", "unit": "atom-lag updates/s",
             "value": total_upd / t["acf_kernels"],
             "roofline": dict(fp32_roof(FLOP_PER_ACF * upd_per_step, t["acf_kernels"]),
                              kernel="acf_band_kernel (+prefix, windows)",
                              hbm=hbm_roof(12.0 * shard * n_frames, t["acf_kernels"]))},
            {"metric": "unwrap_atom_frames_per_s", "unit": "atom-frames/s",
             "value": total_af / t["unwrap_kernel"],
             "roofline": dict(hbm_roof(24.0 * shard * n_frames, t["unwrap_kernel"]),
                              kernel="unwrap_kernel", algorithmic="24 B per atom-frame")},
            {"metric": "ionic_current_atom_frames_per_s", "unit": "atom-frames/s",
             "value": total_af / t["ionic_kernel"],
             "roofline": dict(hbm_roof(12.0 * shard * n_frames, t["ionic_kernel"]),
                              kernel="ionic_current_kernel", algorithmic="12 B per atom-frame")},
        ],
    }
    # the correlation kernels over the lag range: fraction of BOTH rooflines at every lag count
    # (HBM: 12 B per atom-frame read once; FP32: 9 FLOP per MSD update, 6 per ACF update).  Kernel
    # by lag count: <= 4 streaming, 5..128 register-window, beyond: register ring / band Gram
    def _kernel_name(kind, n):
        if n <= 4:
            return f"{kind}_stream_kernel"
        if n <= 128:
            return f"{kind}_rw_kernel"
        return "msd_dense_kernel" if kind == "msd" else "acf_band_kernel"

    line["lag_sweep"] = [
        {"kernel": _kernel_name(kind, n) + (" (+prefix, windows)" if kind == "acf" else ""),
         "data_range": n, "updates_per_s": upd / t,
         "GBps": 12.0 * shard * n_frames / t * 1e-9,
         "frac_of_hbm_peak": 12.0 * shard * n_frames / t * 1e-9 / hbm_peak,
         "frac_of_fp32_peak": flop * upd / t * 1e-12 / fp32_peak}
        for n, t_msd, t_acf, upd in probes
        for kind, t, flop in (("msd", t_msd, FLOP_PER_MSD), ("acf", t_acf, FLOP_PER_ACF))]
    line["hbm_regime"] = [e for e in line["lag_sweep"] if e["data_range"] <= 8]
    if strong is not None:
        line["strong"] = strong
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["rdf"][0], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": cpu["rdf"][1], "host_cores_available": cores}
        line["secondary"][0]["cpu_baseline"] = {"value": cpu["msd"][0], "cores": 1,
                                                "kind": "port", "sample": cpu["msd"][1]}
        line["secondary"][1]["cpu_baseline"] = {"value": cpu["acf"][0], "cores": 1,
                                                "kind": "port", "sample": cpu["acf"][1]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_strong_leg(args, world, rank, dev, barrier):
    """Time-to-solution of the FIXED C5 problem at `world` ranks, through the public calculator
    API and the product's own sharding (no local_only): every rank opens the same project; the
    store keeps this rank's atom block in page-locked host memory; the RDF shards frames (one
    all-to-all of the sampled frames + one all-reduce of the histograms), unwrap / MSD / ACF
    run on the rank's atom block (all-reduce of the series).  Inputs are host resident, results
    land on the host; fits and post-processing are inside the timed region."""
    import torch

    from lammps_analysis_b200.config import config as mdk_config
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import device_fluid
    import tempfile

    small = args.small
    n_sp = args.strong_atoms or (N_SPECIES_ATOMS if not small else 16_000)
    n_frames = args.strong_frames or (N_FRAMES if not small else 600)
    N = DATA_RANGE if not small else 100
    n_cfg = args.strong_rdf_configs if not small else 8
    box_l = BOX if not small else (2 * n_sp / DENSITY) ** (1 / 3)
    box = np.array([box_l] * 3)
    mdk_config.planner_memory_bytes = 60e9
    project = Project("strong", storage_path=tempfile.mkdtemp(prefix="mdk_strong_"),
                      persist=False)          # sharded whenever world > 1
    exp = project.add_experiment("melt", timestep=0.002, temperature=300.0, units="real")
    # the same global trajectory for every N: atoms are generated in fixed chunks of 1/8 species
    # (seeded by species and chunk), each rank makes the chunks of the atom block it owns
    chunk = n_sp // 8
    t_setup = time.perf_counter()
    from lammps_analysis_b200.distributed import shard_atoms
    from lammps_analysis_b200.file_io import BlockInput

    def blocks():
        for si, sp in enumerate(("A", "B")):
            lo, hi = shard_atoms(0, n_sp, rank, world) if world > 1 else (0, n_sp)
            for c0 in range(lo, hi, chunk):
                c1 = min(hi, c0 + chunk)
                seed = 4000 + 100 * si + c0 // chunk
                p = device_fluid(c1 - c0, n_frames, box_l, seed, dev, sigma_step=0.4)
                yield sp, "Positions", (c0, c1), p.cpu().numpy()
                del p
                gen = torch.Generator(device=dev)
                gen.manual_seed(seed + 50)
                v = torch.randn(c1 - c0, n_frames, 3, device=dev, generator=gen)
                yield sp, "Velocities", (c0, c1), v.cpu().numpy()
                del v

    props = {"Positions": 3, "Velocities": 3}
    exp.add_data(BlockInput(n_frames, {"A": (n_sp, props), "B": (n_sp, props)}, box, blocks))
    torch.cuda.empty_cache()
    t_setup = time.perf_counter() - t_setup

    def reset():
        exp.store.invalidate()
        for sp in ("A", "B"):
            if exp.store.check_existence(f"{sp}/Unwrapped_Positions"):
                exp.store.remove(f"{sp}/Unwrapped_Positions")
        exp.version += 1            # defeat the result cache
        barrier()

    def timed_pass(n_configs):
        reset()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        h2d0 = exp.store.h2d_bytes
        w0 = time.perf_counter()
        marks[0].record()
        rdf = exp.run.RadialDistributionFunction(number_of_configurations=n_configs, plot=False)
        marks[1].record()
        from lammps_analysis_b200 import trace as _tr
        _tr.event("RDF done / Einstein call starts")
        ein = exp.run.EinsteinDiffusionCoefficients(data_range=N, plot=False)
        marks[2].record()
        gk = exp.run.GreenKuboDiffusionCoefficients(data_range=N, plot=False)
        exp.store.flush()           # write-back of the unwrapped positions has landed
        marks[3].record()
        from lammps_analysis_b200 import trace
        trace.mark("strong pass: flushed")
        if rank == 0:
            trace.dump()
        barrier()
        wall = time.perf_counter() - w0
        t = [marks[i].elapsed_time(marks[i + 1]) * 1e-3 for i in range(3)]
        tt = torch.tensor(t + [wall], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.cpu().tolist(), rdf, ein, gk, exp.store.h2d_bytes - h2d0

    # warm-up: kernels, NCCL channels, and the page-locked block of the unwrapped positions
    # (its first allocation costs ~0.4 s per GB on these hosts; the store's pool reuses it)
    (_, c_ein, c_gk, _), _, _, _, _ = timed_pass(2)
    prof = None
    if args.strong_profile and rank == 0:      # host-side profile of the timed pass (debug aid)
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    (t_rdf, t_ein, t_gk, wall), rdf, ein, gk, h2d = timed_pass(n_cfg)
    if prof is not None:
        prof.disable()
        prof.dump_stats(args.strong_profile)
    n_tot = 2 * n_sp
    W = n_frames - N                # one frame batch, correlation_time 1
    pairs = n_cfg * n_tot * (n_tot - 1) // 2
    updates = 2 * W * n_sp * N      # per calculator, both species
    if rank != 0:
        return None
    return {
        "problem": f"C5 fixed: 2 x {n_sp} atoms x {n_frames} frames host-resident; "
                   f"RadialDistributionFunction(number_of_configurations={n_cfg}) + "
                   f"EinsteinDiffusionCoefficients(data_range={N}) (runs CoordinateUnwrapper) + "
                   f"GreenKuboDiffusionCoefficients(data_range={N}), public calls, fits included",
        "n_gpus": world, "sharded_store": bool(project.sharded),
        "t_s": t_rdf + t_ein + t_gk, "t_rdf_s": t_rdf, "t_einstein_s": t_ein,
        "t_green_kubo_s": t_gk, "wall_s": wall,
        "rdf_pair_distances_per_s": pairs / t_rdf,
        "dynamics_atom_lag_updates_per_s": 2 * updates / (t_ein + t_gk),
        "h2d_bytes_this_rank": h2d, "setup_s": t_setup,
        # the dynamics part is bound by the host link: 48 GB up (positions, velocities) and 24 GB
        # down (unwrapped positions) over all ranks in t_einstein_s + t_green_kubo_s; compare
        # with scripts/hostlink_probe.py (profiles/r02_hostlink_probe_*): 55 / 57 GB/s H2D / D2H
        # for one rank, 227 / 115 GB/s aggregate for eight on the test box
        "dynamics_host_link_GBps_aggregate": {
            "h2d": 2 * 2 * n_sp * n_frames * 12 / (t_ein + t_gk) * 1e-9,
            "d2h": 2 * n_sp * n_frames * 12 / (t_ein + t_gk) * 1e-9},
        "t_dynamics_first_pass_s": c_ein + c_gk,
        "D_A": ein["A"]["diffusion_coefficient"], "gk_D_A": gk["A"]["diffusion_coefficient"][0],
        "rdf_checksum": float(np.nansum(np.array(rdf["A_B"]["y"])[1:])),
        "note": "time-to-solution is the max over ranks (CUDA events around the public calls); "
                "speed-up at N GPUs = t_s(1) / t_s(N) of the driver's per-N runs",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="reduced sizes (debug only; not a bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the fixed-problem (strong scaling) leg")
    ap.add_argument("--strong-frames", type=int, default=None,
                    help="frames of the strong-scaling trajectory (default: the full 2,000)")
    ap.add_argument("--strong-rdf-configs", type=int, default=64)
    ap.add_argument("--strong-atoms", type=int, default=None,
                    help="atoms per species of the strong-scaling system (default: 500,000)")
    ap.add_argument("--strong-profile", default=None,
                    help="write a cProfile of rank 0's timed strong-scaling pass to this file")
    ap.add_argument("--strong-only", action="store_true",
                    help="run only the strong-scaling leg and print its record (debug aid)")
    args = ap.parse_args()

    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
