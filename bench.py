#!/usr/bin/env python
"""bench.py — MCL sensor-update benchmark (BASELINE.json metric: particle-point evaluations/s + update latency).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4|c5] [--impl b200|reference|reference-cuda]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one full sensor update on one scan: scan preparation, particle x point TSDF evaluation, weight normalisation +
weighted mean pose + CDF, systematic resampling (and, at N > 1 GPUs, the exchange of weights and resampled particles).
Workloads are the BASELINE.json configs built by tsdf_localization_b200/synthetic.py (seeds fixed):
  c3 (default)  OS1-128 scan (131,072 points) x 65,536 particles, box room 20x20x5 m @ 5 cm — the config the north star's
                <5 ms target is quoted on; STRONG scaling: the 65,536 particles are sharded over the N ranks
  c2            VLP-16 scan (30,000 points) x 5,000 particles            c1   1,024 points x 500 particles (the parity config)
  c4            1,048,576 uniform particles x reduced OS1-128 scan, 1.4 GB multi-room map (global localisation)
  c5            262,144 tracking particles x OS1-128 scan, the same 1.4 GB map (> L2)
`value`  : whole-job particle-point evaluations/s, inputs already resident in HBM, CUDA events around each step on the
           launching stream, L2 flushed (256 MiB memset) between steps outside the timed events, max over ranks.
`e2e`    : the same update through the reference-facing host-buffer C-ABI calls, host<->device copies inside the timed region:
           tsdfloc_sensor_update + tsdfloc_resample_systematic at N = 1; tsdfloc_multi_sensor_update +
           tsdfloc_multi_resample_systematic issued by rank 0 over all N devices at N > 1 (the form in which the reference's
           single-process node binds the library).
`roofline`: the evaluation kernel k_eval alone (tsdfloc_last_eval_ms: CUDA events on its stream), algorithmic bytes =
           8 B per particle-point evaluation (4 B brick-table entry + 4 B voxel) over the measured HBM copy peak;
           `bound` names the resource that actually binds per workload (profiles/k_eval_ncu.json holds the ncu figures).
`cpu_baseline` / --impl reference: the UNMODIFIED reference CPU/OpenMP evaluator + SystematicResampler (oracle/_ref,
           compiled from /root/reference in the build container) on this box's host cores, on a bounded particle sample.
--impl reference-cuda: the reference's OWN CUDA evaluator (src/cuda/*.cu compiled unmodified for sm_100a into
           oracle/_ref/libtsdf_ref_cuda.so) + its CPU SystematicResampler, the full workload, host buffers in and out —
           what the node does today on the same GPU. The b200 line carries its figure as `reference_cuda_ms_per_update`.
The reference arms build their maps and inputs without loading libtsdfloc.so.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle_point_evals_per_s"
UNIT = "evals/s"
ALG_BYTES_PER_EVAL = 8          # 4 B brick-table entry + 4 B voxel
L2_FLUSH_BYTES = 256 << 20

PARTICLES_OVERRIDE = 0      # --particles: development runs of the box-room workloads at another particle count

WORKLOADS = {
    "c3": dict(map="box", scan="os1-128", particles=65536, desc="OS1-128 scan (131,072 pts) x 65,536 particles, box room 20x20x5 m @ 5 cm"),
    "c2": dict(map="box", scan="vlp16", particles=5000, desc="VLP-16 scan (30,000 pts) x 5,000 particles, box room 20x20x5 m @ 5 cm"),
    "c1": dict(map="box", scan="vlp16", particles=500, n_points=1024, desc="1,024 pts x 500 particles, box room 20x20x5 m @ 5 cm"),
    "c4": dict(map="rooms", desc="global localisation: 1,048,576 uniform particles x OS1-128 scan after 0.256 m ring-aware reduction, "
                                 "multi-room 100x100x10 m map (1.4 GB, > L2)"),
    "c5": dict(map="rooms", desc="262,144 tracking particles x OS1-128 scan (131,072 pts), multi-room 100x100x10 m map (1.4 GB, > L2)"),
}
# what binds the evaluation kernel per workload (ncu: profiles/k_eval_ncu.json): DRAM traffic is a small fraction of the
# algorithmic bytes wherever the touched bricks fit in L2 — then the SM's FP32 pipe / issue slots and the L2->L1 sector
# stream bind; only the global-localisation config gathers all over a map 11x the L2
BOUND = {"c1": "launch latency (one partial wave)", "c2": "l2/fp32-pipe", "c3": "l2/fp32-pipe", "c4": "hbm", "c5": "l2/fp32-pipe"}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- workloads: inputs are numpy only; maps are built per arm -----------------------------------------------------------------

def workload_inputs(name: str):
    """(particles [N,7], points [P,3], tf[16]) — pure numpy generators, no native library involved."""
    import common
    from tsdf_localization_b200 import synthetic as syn
    w = WORKLOADS[name]
    if name == "c4":
        ps, pts, _ = common.config_c4()
    elif name == "c5":
        ps, pts, _ = common.config_c5()
    else:
        pts, _ = syn.make_scan(w["scan"], syn.GT_POSE, n_points=w.get("n_points"))
        ps = syn.tracking_particles(PARTICLES_OVERRIDE or w["particles"], syn.GT_POSE)
    return np.ascontiguousarray(ps), np.ascontiguousarray(pts), syn.IDENTITY_TF


def product_map(name: str):
    """The product's host map object (libtsdfloc.so's own map builder)."""
    import common
    return common.grid_rooms() if WORKLOADS[name]["map"] == "rooms" else common.box_room()[1]


class _RefMapGeometry:
    """What synthetic.grid_rooms_arrays needs from a map object: coef()."""

    def __init__(self, ref, handle):
        self.ref, self.handle = ref, handle

    def coef(self):
        return self.ref.map_coef(self.handle)


def reference_map(ref, name: str):
    """The same map inside the UNMODIFIED reference's CudaSubVoxelMap, built without libtsdfloc.so: the likelihood values come
    from the oracle's restatement of createTSDFMap's transform, the bricks from setData (box room) or — for the 1.4 GB map,
    which is generated as arrays — through the map's own raw-array accessors."""
    from oracle_lib import Oracle
    from tsdf_localization_b200 import synthetic as syn
    o = Oracle()
    value_fn = lambda mm, sigma=syn.SIGMA: float(o.lib.oracle_likelihood_value(C.c_float(mm), C.c_float(sigma)))   # noqa: E731
    init = float(o.lib.oracle_likelihood_init(C.c_float(syn.SIGMA)))
    if WORKLOADS[name]["map"] == "box":
        spec = syn.box_room_map(value_fn, init)
        rm = ref.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
        if not rm or ref.map_set_data(rm, spec.cells) != 0:
            raise RuntimeError("reference map build failed")
        return rm
    holder = {}

    def make(mn, mx, res, init_value):
        holder["rm"] = ref.map_create(mn, mx, res, init_value)
        if not holder["rm"]:
            raise RuntimeError("reference map build failed")
        return _RefMapGeometry(ref, holder["rm"])

    _, occ, data = syn.grid_rooms_arrays(make, value_fn, init)
    if ref.map_adopt_arrays(holder["rm"], occ, data) != 0:
        raise RuntimeError("reference map build failed: " + ref.last_error())
    return holder["rm"]


# ---- clocks sampler ---------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- reference arms / cpu baseline --------------------------------------------------------------------------------------

def pick_ref_threads():
    """The reference hard-codes OMP_THREADS = 8 (util/constant.h:4); oracle/_ref holds variants rebuilt with the constant
    shadowed. Use the largest variant that does not exceed this host's cores."""
    from oracle_lib import ref_lib_path
    cores = os.cpu_count() or 8
    best = None
    for t in (8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256):
        path = ref_lib_path(None if t == 8 else t)
        if t <= max(cores, 8) and path.exists():
            best = t
    return best


def _ref_eval_call(ref, ev, ps, pts, tf, use_cuda: bool):
    """One ref_evaluate call on prepared arrays (the harness copies them into the reference's std::vectors, like the node's
    own callback fills them): returns the particle array carrying the normalised weights."""
    out = np.array(ps, dtype=np.float32, copy=True, order="C")
    pose = np.zeros(7, dtype=np.float64)
    rc = ref.lib.ref_evaluate(ev, out.ctypes.data_as(C.c_void_p), out.shape[0], pts.ctypes.data_as(C.c_void_p), pts.shape[0],
                              tf.ctypes.data_as(C.c_void_p), 1 if use_cuda else 0, pose.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("reference evaluate failed: " + ref.lib.ref_last_error().decode())
    return out


def reference_runner(name, ps, pts, tf, sample: int):
    """Returns (run_once() -> seconds for evaluate + resample on the sample, cores, description)."""
    from oracle_lib import Ref, ref_available
    if not ref_available():
        raise RuntimeError("oracle/_ref not built: run `make -C oracle` in the container that has /root/reference")
    t = pick_ref_threads()
    ref = Ref(None if t == 8 else t)
    rm = reference_map(ref, name)
    ev = ref.eval_create(rm)
    sub = np.ascontiguousarray(ps[:sample])
    tfa = np.ascontiguousarray(tf, dtype=np.float32)

    def run_once():
        t0 = time.perf_counter()
        out = _ref_eval_call(ref, ev, sub, pts, tfa, False)
        ref.systematic_resample(out, 1)
        return time.perf_counter() - t0

    return run_once, ref.omp_threads(), f"first {sample} of {len(ps)} particles x full {len(pts)}-point scan"


def cpu_sample_size(n_particles: int, n_points: int, evals: float = 3.0e8) -> int:
    """Particles of the bounded CPU sample: ~`evals` particle-point evaluations per run (3e8 = about 0.3 s on 16 cores),
    capped by the workload."""
    return int(max(16, min(n_particles, evals // max(n_points, 1))))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        ps, pts, tf = workload_inputs(args.workload)
        sample = cpu_sample_size(len(ps), len(pts))
        run_once, cores, what = reference_runner(args.workload, ps, pts, tf, sample)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0]}))
        return
    for _ in range(args.warmup):
        run_once()
    times = [run_once() for _ in range(args.steps)]
    total = sum(times)
    value = sample * len(pts) * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + WORKLOADS[args.workload]["desc"], "particles": len(ps), "points": len(pts),
                   "sample": what, "step": "TSDFEvaluator::evaluate(use_cuda=false) + SystematicResampler::resample on the sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_cuda_runner(name, ps, pts, tf):
    """run_once() -> seconds of one full update by the reference's own CUDA evaluator: TSDFEvaluator::evaluate(use_cuda=true)
    (src/cuda/cuda_evaluator.cu:118-428, its own mallocs, copies and syncs) + SystematicResampler::resample on the CPU, host
    buffers in and out — the methodology of src/num_particles_eval.cpp:222-257."""
    from oracle_lib import Ref, ref_cuda_path
    if not ref_cuda_path().exists():
        raise RuntimeError("oracle/_ref/libtsdf_ref_cuda.so not built (needs /root/reference and nvcc at build time)")
    ref = Ref(cuda=True)
    rm = reference_map(ref, name)
    ev = ref.eval_create(rm)
    if not ev:
        raise RuntimeError("reference CudaEvaluator: " + ref.last_error())
    tfa = np.ascontiguousarray(tf, dtype=np.float32)

    def run_once():
        t0 = time.perf_counter()
        out = _ref_eval_call(ref, ev, ps, pts, tfa, True)
        t1 = time.perf_counter()
        ref.systematic_resample(out, 1)
        return time.perf_counter() - t0, t1 - t0

    return run_once


def run_reference_cuda_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        ps, pts, tf = workload_inputs(args.workload)
        run_once = reference_cuda_runner(args.workload, ps, pts, tf)
        for _ in range(max(1, args.warmup)):
            run_once()
        rows = [run_once() for _ in range(args.steps)]
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference-cuda", "unavailable": str(e).splitlines()[0]}))
        return
    total = sum(r[0] for r in rows)
    value = len(ps) * len(pts) * len(rows) / total
    line = {
        "impl": "reference-cuda", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(rows), "evaluate_ms": 1e3 * sum(r[1] for r in rows) / len(rows),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + WORKLOADS[args.workload]["desc"], "particles": len(ps), "points": len(pts),
                   "step": "the reference's own CudaEvaluator::evaluate (src/cuda, unmodified, sm_100a) incl. its copies + "
                           "SystematicResampler::resample on the CPU; full workload, host buffers"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": len(ps) * 52 + len(pts) * 12, "d2h_bytes_per_step": len(ps) * 4},
    }
    print(json.dumps(line))


# ---- B200 arm ----------------------------------------------------------------------------------------------------------

class StdoutGuard:
    """Keeps stdout to the ONE JSON line: file descriptor 1 is pointed at stderr while the run is in progress (NCCL and other
    native libraries print banners straight to fd 1), and the line is written to the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str) -> None:
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def close(self) -> None:
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def run_b200_arm(args):
    guard = StdoutGuard()
    try:
        _run_b200_arm(args, guard)
    finally:
        guard.close()


def _ncu_figures(workload: str):
    path = ROOT / "profiles" / "k_eval_ncu.json"
    if not path.exists():
        return {}
    try:
        return json.loads(path.read_text()).get(workload, {})
    except Exception:  # noqa: BLE001
        return {}


def _run_b200_arm(args, guard):
    import torch
    import torch.distributed as dist

    from tsdf_localization_b200 import CudaEvaluator, MultiGpuEvaluator, capi
    from tsdf_localization_b200.dist import GpuStages, ShardedSensorUpdate, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py ...")
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the sensor update has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_group = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side barrier: ranks waiting on it keep their GPUs idle

    ps, pts, tf = workload_inputs(args.workload)
    m = product_map(args.workload)
    n, p = len(ps), len(pts)
    ev = CudaEvaluator(m, device=local)
    if args.registers:
        ev.tune(capi.TUNE_EVAL_REGISTERS, args.registers)
    if not args.graphs:
        ev.tune(capi.TUNE_GRAPHS, 0)
    if args.chunks:
        ev.tune(capi.TUNE_EVAL_CHUNKS, args.chunks)
    if args.pairing:
        ev.tune(capi.TUNE_EVAL_PAIRING, args.pairing)
    lib = capi.load_library()
    stages = GpuStages(ev)
    fused = {"auto": None, "nccl": False, "fused": True}[args.transport]
    upd = ShardedSensorUpdate(stages, world=world, rank=rank, device=dev, max_particles=n, fused=fused)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    d_ps = torch.from_numpy(ps).to(dev)
    d_pts = torch.from_numpy(pts).to(dev)
    h_ps = torch.from_numpy(ps).pin_memory()
    h_pts = torch.from_numpy(pts).pin_memory()
    u0 = 0.37 / n
    _, _, n_local = shard(n, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        upd.set_scan(d_pts)
        return upd.step(d_ps, n, tf, u0)

    # e2e: the reference-facing host-buffer C-ABI calls. N = 1: tsdfloc_sensor_update + tsdfloc_resample_systematic.
    # N > 1: rank 0 alone drives all N devices through tsdfloc_multi_* (what the single-process node binds); the other
    # ranks wait on a host-side barrier with idle GPUs.
    cap = n + n // 8 + 64
    h_out = torch.empty((cap, 7), dtype=torch.float32).pin_memory()
    e2e_ps = h_ps.clone().pin_memory()
    tfc = (C.c_float * 16)(*[float(v) for v in tf])
    multi, multi_err = None, None
    if world > 1 and rank == 0:
        try:
            multi = MultiGpuEvaluator(m, devices=list(range(world)))
        except Exception as e:  # noqa: BLE001  (e.g. the launcher restricted this rank to one visible device)
            multi_err = str(e).splitlines()[0]

    def e2e_step():
        mean = (C.c_float * 6)()
        n_out = C.c_uint64(0)
        if world == 1:
            capi.check(lib, ev.ctx, lib.tsdfloc_sensor_update(ev.ctx, C.c_void_p(e2e_ps.data_ptr()), n, C.c_void_p(h_pts.data_ptr()), p, tfc, mean))
            capi.check(lib, ev.ctx, lib.tsdfloc_resample_systematic(ev.ctx, C.c_float(u0), C.c_void_p(h_out.data_ptr()), cap, C.byref(n_out), None))
        else:
            multi._check(lib.tsdfloc_multi_sensor_update(multi._m, C.c_void_p(e2e_ps.data_ptr()), n, C.c_void_p(h_pts.data_ptr()), p, tfc, mean))
            multi._check(lib.tsdfloc_multi_resample_systematic(multi._m, C.c_float(u0), C.c_void_p(h_out.data_ptr()), cap, C.byref(n_out)))
        return int(n_out.value)

    def e2e_step_per_process():
        d_ps.copy_(h_ps, non_blocking=True)
        d_pts.copy_(h_pts, non_blocking=True)
        upd.set_scan(d_pts)
        out, mean, n_out, _ = upd.step(d_ps, n, tf, u0)
        h_out[:n_out].copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return n_out

    with torch.cuda.stream(stream):
        # ---- device-resident timing -------------------------------------------------------------------------------------
        for _ in range(args.warmup):
            device_step()
        barrier()
        launches0 = stages.kernel_launches()
        sampler = ClockSampler(local) if rank == 0 else None
        step_ms, eval_ms = [], []
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()                                   # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out, _, n_out, wsum = device_step()
            e1.record(stream)
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            eval_ms.append(stages.last_eval_ms())
        barrier()
        wall = time.perf_counter() - wall0
        launches = stages.kernel_launches() - launches0
        clocks = sampler.stop() if sampler else None
        # what the update computed, for rank-count invariance: every rank holds the full vectors; rank 0 reports them
        weights_sha = sha(d_ps[:n, 6].cpu().numpy())
        resampled_sha = sha(out[:n_out].cpu().numpy())

        # ---- per-process end-to-end (pinned copies around the sharded update on every rank) -------------------------------
        pp_ms = []
        if world > 1:
            for _ in range(max(1, min(args.warmup, 3))):
                e2e_step_per_process()
            barrier()
            for _ in range(args.steps):
                flush.zero_()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                e2e_step_per_process()
                pp_ms.append(1e3 * (time.perf_counter() - t0))
            barrier()

        # ---- end-to-end through the host-buffer C ABI ---------------------------------------------------------------------
        e2e_ms = []
        e2e_sha = None
        if world == 1 or (rank == 0 and multi is not None):
            for _ in range(max(1, min(args.warmup, 3))):
                e2e_step()
            for _ in range(args.steps):
                flush.zero_()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                m_out = e2e_step()
                e2e_ms.append(1e3 * (time.perf_counter() - t0))
            e2e_sha = (sha(e2e_ps.numpy()[:, 6]), sha(h_out.numpy()[:m_out]))
        if world > 1:
            dist.barrier(group=host_group)

    t = torch.tensor([sum(step_ms), sum(pp_ms), float(np.mean(eval_ms)), wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, pp_total_ms, eval_mean_ms, wall_ms = (float(v) for v in t.cpu())
    K = args.steps
    ms_per_step = total_ms / K
    value = n * p / (ms_per_step * 1e-3)

    if rank == 0:
        via_multi = world > 1 and multi is not None
        if world > 1 and not via_multi:      # rank 0 cannot see the other devices: report the per-process figure instead
            e2e_ms, e2e_sha = [pp_total_ms / K] * K, (weights_sha, resampled_sha)
        e2e_total_ms = sum(e2e_ms)
        e2e_value = n * p / (e2e_total_ms / K * 1e-3)
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = ALG_BYTES_PER_EVAL * n_local * p / (eval_mean_ms * 1e-3) / 1e9
        ncu = _ncu_figures(args.workload)
        # complementary view: fraction of the chip's MEASURED ceiling for independent 4 B gathers out of an
        # L2-resident array when the 32 lanes of a request spread over as many sectors as the kernel's do
        # (scripts/probe_gather.py -> profiles/r01c_probe_gather.jsonl; the kernel issues 2 gathers per evaluation)
        gather_view = None
        ppath = ROOT / "profiles" / "r01c_probe_gather.jsonl"
        if ppath.exists():
            try:
                rows = [json.loads(l) for l in ppath.read_text().splitlines() if l.strip()]
                spr = ncu.get("sectors_per_request", 13.0)
                cands = [r for r in rows if not isinstance(r["lanes_share_sectors"], str)]
                row = min(cands, key=lambda r: abs(float(r["lanes_share_sectors"]) - spr))
                mine = 2.0 * n_local * p / (eval_mean_ms * 1e-3)
                gather_view = {"achieved_gathers_per_s": mine, "probe_gathers_per_s": row["gathers_per_s"],
                               "probe_sectors_per_request": row["lanes_share_sectors"], "frac": mine / row["gathers_per_s"],
                               "source": "profiles/r01c_probe_gather.jsonl (k_probe_gather at the kernel's sectors per request)"}
            except Exception:  # noqa: BLE001
                gather_view = None
        roofline = {"bound": BOUND[args.workload], "kernel": "k_eval", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu.get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel_ms": eval_mean_ms,
                    "algorithmic_bytes_per_launch": ALG_BYTES_PER_EVAL * n_local * p,
                    "fma_pipe_pct": ncu.get("fma_pipe_pct"), "lts_pct": ncu.get("lts_pct"), "issue_active_pct": ncu.get("issue_active_pct"),
                    "ncu_source": ncu.get("source"), "l2_gather_view": gather_view,
                    "note": "frac = 8 B per particle-point evaluation (4 B brick-table entry + 4 B voxel) over the measured HBM copy peak, as the "
                            "contract defines it; `bound` names what binds: with the touched bricks L2-resident DRAM traffic is ~1 % of the "
                            "algorithmic bytes and the SM FP32 pipe / issue slots + the L2->L1 sector stream limit the kernel (DESIGN.md)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sample = cpu_sample_size(n, p, 2.0e9)      # ~2 s per run on 16 cores; 1 warm-up + best of 5 = ~12 s of CPU work
                run_once, cores, what = reference_runner(args.workload, ps, pts, tf, sample)
                run_once()
                best = min(run_once() for _ in range(5))
                cpu = {"value": sample * p / best, "unit": UNIT, "cores": cores, "kind": "reference", "sample": what + ", best of 5",
                       "ms_per_update_extrapolated": 1e3 * best * n / sample}
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: " + str(e).splitlines()[0]}
        ref_cuda_ms = None
        if world == 1 and not args.no_reference_cuda:
            try:
                torch.cuda.synchronize(dev)
                run_once = reference_cuda_runner(args.workload, ps, pts, tf)
                run_once()
                ref_cuda_ms = 1e3 * min(run_once()[0] for _ in range(3))
            except Exception as e:  # noqa: BLE001
                ref_cuda_ms = "unavailable: " + str(e).splitlines()[0]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload + ": " + WORKLOADS[args.workload]["desc"] + (" [--particles override]" if PARTICLES_OVERRIDE else ""),
                       "particles": n, "points": p, "particles_per_gpu": n_local, "map_mb": m.dataBytes() / 1e6, "sharding": f"particles/{world}, map replicated",
                       "step": "scan prep + eval + normalise/mean/CDF + systematic resample" + (
                           "" if world == 1 else " + 2 NCCL all-gathers" if upd.transport == "all_gather"
                           else "; eval and draw kernels store into all peers' symmetric-memory buffers over NVLink (no collective), 3 signal barriers"),
                       "transport": upd.transport,
                       "l2": "flushed (256 MiB memset) between timed steps, outside the CUDA-event pairs",
                       "n_out": int(n_out), "weight_sum": wsum,
                       # rank-count invariance: these four hashes must be identical at every --gpus N
                       "weights_sha256": weights_sha, "resampled_sha256": resampled_sha,
                       "e2e_weights_sha256": e2e_sha[0], "e2e_resampled_sha256": e2e_sha[1]},
            "roofline": roofline, "cpu_baseline": cpu, "reference_cuda_ms_per_update": ref_cuda_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_total_ms / K,
                    "h2d_bytes_per_step": world * (n * 28 + p * 12), "d2h_bytes_per_step": int(n_out) * 28 + n * 28 + 24,
                    "api": "tsdfloc_sensor_update + tsdfloc_resample_systematic (host buffers)" if world == 1
                           else f"tsdfloc_multi_sensor_update + tsdfloc_multi_resample_systematic from rank 0 over {world} devices (host buffers)"
                           if via_multi else f"pinned H2D + ShardedSensorUpdate.step + pinned D2H per rank (tsdfloc_multi unavailable: {multi_err})",
                    "per_process_ms_per_step": (pp_total_ms / K) if world > 1 else None},
            "gpu_launches": int(launches), "graphs": dict(zip(("recordings", "replays", "note"), ev.graph_stats())), "clocks": clocks, "wall_ms_per_step_incl_flush": wall_ms / K,
        }
        guard.emit(json.dumps(line))
    if multi is not None:
        multi.close()
    ev.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("b200", "reference", "reference-cuda"), default="b200")
    ap.add_argument("--workload", choices=tuple(WORKLOADS), default="c3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--transport", choices=("auto", "nccl", "fused"), default="auto",
                    help="multi-GPU exchange: fused = kernels store straight into the peers' buffers (default when available); nccl = all-gathers")
    ap.add_argument("--particles", type=int, default=0, help="development: override the particle count of c1-c3 (not a BASELINE config)")
    ap.add_argument("--registers", type=int, default=0, choices=(0, 1, 2),
                    help="development: force the evaluation kernel's register budget (1: 64, 2: 128; 0 = the library's own rule)")
    ap.add_argument("--pairing", type=int, default=0, choices=(0, 1, 2),
                    help="development: force the evaluation kernel's pairing (1: two particles per warp, 2: two points per lane)")
    ap.add_argument("--chunks", type=int, default=0,
                    help="development: scan chunks per particle in the evaluation kernel (1 = whole scans; 0 = the library's own rule)")
    ap.add_argument("--graphs", type=int, default=1, choices=(0, 1),
                    help="development: 0 = launch every update kernel by kernel (A/B against the steady-state CUDA graphs)")
    args = ap.parse_args()
    global PARTICLES_OVERRIDE
    PARTICLES_OVERRIDE = args.particles
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
