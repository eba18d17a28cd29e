#!/usr/bin/env python
"""bench.py -- throughput of the spectral-conversion hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path fbf|traj|dtw]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); rank 0 prints ONE JSON line.

  --path fbf  (default; BASELINE.json configs[1], "C1")   converted frames/s of vc(::GMMMap) over
              1,000,000 synthetic 24-dim frames with a 64-mixture full-covariance joint GMM per GPU
              (weak scaling, no data-path collective).  The default run also carries the other
              paths' complete sub-lines under "other_paths".
  --path traj (configs[4], "C4")  128-mixture trajectory conversion of ONE batch of 8192 DISTINCT
              utterances x 500 frames, sharded by utterance over the ranks with shard.shard_ragged
              (strong scaling); the optional NCCL gather of the results is timed separately.
  --path dtw  (configs[3], "C3")  DTW(fstep=0, bstep=2) of 1000 pairs (~600 x 600, 24-dim) per GPU.

--impl reference times the CPU restatement of the Julia reference (the oracle port; Julia itself is
not installable in this image) on the host cores, for the same path and config.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

M_MIX, DIM, FRAMES = 64, 24, 1_000_000
F_FBF = 4 * M_MIX * DIM * DIM + 2 * M_MIX * DIM          # 150,528 flop/frame
B_TRAJ = 48 * 24 * 24 + 24 * 24                          # 28,224 B/frame (Ds = 24)
B_DTW = 17.0                                             # B/cell in the two-pass form
DP_OPS_PER_CELL = 3 * 24 + 7                             # FP64 operations per DTW cell (D = 24)
C4_UTT, C4_FRAMES, C4_MIX = 8192, 500, 128

METRICS = {
    "fbf": "converted frames/sec (GMM frame-by-frame, C1)",
    "traj": "converted frames/sec (trajectory, C4)",
    "dtw": "DTW cells/sec (C3)",
}
UNITS = {"fbf": "frames/s", "traj": "frames/s", "dtw": "cells/s"}


def config_for(path: str, world: int) -> dict:
    """The `config` object of a path; identical in the GPU arm and the --impl reference arm."""
    if path == "fbf":
        return {"workload": "C1: CMU-Arctic-shaped synthetic, 24-dim mcep, 64-mixture full-cov joint GMM, "
                            "frame-by-frame vc() of 1M frames per GPU",
                "frames_per_gpu": FRAMES, "mixtures": M_MIX, "dim": DIM,
                "l2": "per-step input+output = 400 MB per GPU, larger than the 126 MB L2",
                "parallelism": f"frames sharded over {world} GPU(s), no data-path collective"}
    if path == "traj":
        return {"workload": "C4: 128-mixture full-cov trajectory conversion (48-dim static+delta source) of one batch of "
                            "8192 distinct utterances x 500 frames, one chunk per utterance",
                "utterances": C4_UTT, "frames_per_utterance": C4_FRAMES, "mixtures": C4_MIX, "static_dim": 24,
                "chunk_limit": C4_FRAMES,
                "l2": "per-step input+output+factor scratch >> 126 MB L2 on every rank",
                "parallelism": f"utterances sharded over {world} GPU(s) by shard.shard_ragged, no data-path collective; "
                               "optional NCCL gather timed separately"}
    return {"workload": "C3: DTW(fstep=0, bstep=2) of 1000 parallel utterance pairs (~600x600 frames, 24-dim) per GPU",
            "pairs_per_gpu": 1000, "dim": 24, "fstep": 0, "bstep": 2,
            "l2": "the kernel keeps the cost column in registers; templates + sequences + back-pointers = 0.3 GB per step, larger than the 126 MB L2",
            "parallelism": f"pairs sharded over {world} GPU(s), no data-path collective"}


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def host_cores() -> int:
    """Host threads the CPU arm may use.  torchrun exports OMP_NUM_THREADS=1, so the OpenMP default
    is not the machine size; the thread count is passed explicitly (num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ncu_summary(kernel_file: str):
    """Metrics of the dominant kernel from the latest `ncu --set full` summary committed under
    profiles/ (tools/summarize_profiles.py): DRAM traffic per launch and pipe utilisations."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kernel_file}.txt")))
    out = {"traffic": None, "source": None}
    if not files:
        return out
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(files[-1]):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m:
            tot += float(m.group(2)) * unit.get(m.group(3), 1.0)
        m = re.match(r"(sm__pipe_fp64_cycles_active|sm__pipe_tensor_cycles_active|sm__warps_active|launch__grid_size)\S*\s+([0-9.]+)", line)
        if m:
            out[m.group(1)] = float(m.group(2))
    out["traffic"] = tot or None
    out["source"] = os.path.relpath(files[-1], ROOT)
    return out


def measure_tf32_peak(torch):
    """Dense TF32 tensor peak with the driver's method for bf16 (torch.matmul 8192^3, best of 10)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda", dtype=torch.float32)
        b = torch.randn(n, n, device="cuda", dtype=torch.float32)
        for _ in range(3):
            a @ b
        best = 1e30
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); a @ b; e.record(); e.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def bind_near_gpu(torch, local: int) -> dict:
    """Binds this rank to the CPUs local to its GPU (PCIe root / NUMA node) BEFORE any page-locked
    buffer is allocated, so the staging memory is first-touched next to the device.  A container whose
    cpuset excludes those CPUs keeps its affinity (reported)."""
    info = {"bound": False, "cpus": None, "local_cpulist": None}
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        info["local_cpulist"] = txt
        want = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                want.update(range(int(a), int(b) + 1))
            elif part:
                want.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as ex:
        info["error"] = repr(ex)
    return info


def dtw_pairs_for_rank(vcb, rank: int):
    """1000 C3 pairs; rank r aligns its own, distinct set (seed 1003 + r)."""
    return vcb.synth.config_c3(1000, seed=1003 + rank)


# ------------------------------------------------------------------------------------------------
# --impl reference: the CPU restatement of the Julia reference (oracle port) on all host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(path: str, steps: int, warmup: int, cores: int):
    """Times `steps` passes of the oracle over (a bounded sample of) the path's workload.
    Returns (units per second, seconds per step, sample description)."""
    import vcb200 as vcb
    from oracle import oracle as O
    O.build()
    if path == "fbf":
        gm, fm = vcb.synth.config_c1(FRAMES)
        g = O.GMMMap(*gm)
        units, what = FRAMES, (f"all {FRAMES} frames of the C1 workload per step, OpenMP over frames on {cores} threads; C "
                               "restatement of the Julia reference (Julia 0.5 is not installable here and is single-threaded)")
        run = lambda: g.vc(fm, nthreads=cores)                                     # noqa: E731
        warm = lambda: g.vc(np.asfortranarray(fm[:, : FRAMES // 16]), nthreads=cores)   # noqa: E731
    elif path == "traj":
        n = min(C4_UTT, 8 * cores)
        gm = vcb.synth.random_joint_gmm(1004, C4_MIX, 96)
        fm, off = vcb.synth.c4_utterances(gm, np.arange(n), C4_FRAMES)
        g = O.GMMMap(*gm)
        units, what = n * C4_FRAMES, (f"utterances 0..{n - 1} of the C4 batch ({n} x {C4_FRAMES} frames) per step, OpenMP over "
                                      f"utterances on {cores} threads; C restatement of the Julia reference")
        run = lambda: O.vc_traj_batch(g, C4_FRAMES, fm, off, nthreads=cores)       # noqa: E731
        warm = lambda: O.vc_traj_batch(g, C4_FRAMES, np.asfortranarray(fm[:, : off[cores]]), off[: cores + 1], nthreads=cores)  # noqa: E731
    else:
        tm, to, sq, so = dtw_pairs_for_rank(vcb, 0)
        units = float(np.sum(np.diff(to).astype(np.float64) * np.diff(so)))
        what = (f"all 1000 pairs of the C3 workload per step, OpenMP over pairs on {cores} threads; C restatement of the "
                "Julia reference")
        run = lambda: O.dtw_fit_batch(tm, to, sq, so, 0, 2, nthreads=cores)        # noqa: E731
        warm = lambda: O.dtw_fit_batch(tm, to[:65], sq, so[:65], 0, 2, nthreads=cores)  # noqa: E731
    for _ in range(max(warmup, 1)):
        warm()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return units / dt, dt, what


def cpu_reference_arm(args):
    if _env_int("RANK", 0) != 0:
        return
    cores = host_cores()
    value, dt, what = cpu_run(args.path, args.steps, args.warmup, cores)
    unit = UNITS[args.path]
    line = {
        "impl": "reference", "metric": METRICS[args.path], "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.path == "traj" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args.path, args.gpus),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(path: str) -> dict:
    """cpu_baseline object of the GPU arm (rank 0, N = 1): one bounded pass of the oracle."""
    unit = UNITS[path]
    try:
        cores = host_cores()
        if path == "fbf":
            # a bounded slice of the 1M frames keeps the default run short
            import vcb200 as vcb
            from oracle import oracle as O
            O.build()
            gm, fm = vcb.synth.config_c1(FRAMES)
            sample = min(FRAMES, 25_000 * cores)
            og = O.GMMMap(*gm)
            sub = np.asfortranarray(fm[:, :sample])
            og.vc(np.asfortranarray(sub[:, : sample // 10]), nthreads=cores)
            t0 = time.perf_counter(); og.vc(sub, nthreads=cores); dt_all = time.perf_counter() - t0
            n1 = min(sample, 20_000)
            t0 = time.perf_counter(); og.vc(np.asfortranarray(sub[:, :n1])); dt_1 = time.perf_counter() - t0
            return {"value": sample / dt_all, "unit": unit, "cores": cores, "kind": "port",
                    "sample": f"first {sample} frames of the same C1 workload, OpenMP over frames (C restatement of the "
                              "Julia reference; Julia is single-threaded)",
                    "single_thread_value": n1 / dt_1, "single_thread_sample": f"first {n1} frames"}
        value, _, what = cpu_run(path, 1, 1, cores)
        return {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": what}
    except Exception as ex:
        return {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": repr(ex)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import vcb200 as vcb
        self.torch, self.dist, self.vcb, self.args = torch, dist, vcb, args
        self.rank, self.world, self.local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: no CUDA device is visible and there is no CPU path")
        torch.cuda.set_device(self.local)
        self.host = bind_near_gpu(torch, self.local) if self.world > 1 else {"bound": False, "cpus": host_cores()}
        vcb.set_device(self.local)
        vcb.set_kernel_variant(args.variant)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.peaks, self.peak_src = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, steps: int, warmup: int, stage_index=None):
        """W untimed + K timed calls of fn() bracketed by barrier + synchronize; CUDA events on the
        current stream (the stream the library launches on).  Returns (ms per step as the max over
        ranks, launches, mean duration of stage `stage_index` of the library's stage marks)."""
        torch, vcb = self.torch, self.vcb
        for _ in range(warmup):
            fn()
        self.barrier()
        l0 = vcb.launch_count()
        if stage_index is not None:
            vcb.stage_timing(True)      # the library records CUDA events between its stages, per call
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        self.barrier()
        stages = []
        if stage_index is not None:
            stages = vcb.stage_times()      # mean over the timed steps (the last 64 at most)
            vcb.stage_timing(False)
        launches = vcb.launch_count() - l0
        ms = self.max_over_ranks(s.elapsed_time(e) / steps)
        return ms, int(launches), (stages[stage_index] if len(stages) > (stage_index or 0) else None), stages

    def copy_ceiling(self, mb: int = 200, reps: int = 5) -> dict:
        """Platform ceiling of the end-to-end paths, measured live: every rank copies `mb` MB host->device and
        `mb` MB device->host CONCURRENTLY from page-locked memory on two streams, all ranks at once (barrier),
        max over ranks.  GB/s per direction per GPU; on a box whose GPUs share host links it falls with N."""
        if getattr(self, "_ceiling", None) is not None:
            return self._ceiling
        torch = self.torch
        n = mb << 20
        hin = torch.empty(n, dtype=torch.uint8).pin_memory()
        hout = torch.empty(n, dtype=torch.uint8).pin_memory()
        din = torch.empty(n, dtype=torch.uint8, device="cuda")
        dout = torch.empty(n, dtype=torch.uint8, device="cuda")
        s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

        def once():
            with torch.cuda.stream(s0):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s1):
                hout.copy_(dout, non_blocking=True)
        once()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        s0.synchronize(); s1.synchronize()
        dt = self.max_over_ranks((time.perf_counter() - t0) / reps)
        self._ceiling = {"gbs_each_direction_per_gpu": n / dt / 1e9, "gbs_each_direction_all_gpus": self.world * n / dt / 1e9,
                         "how": f"{mb} MB H2D + {mb} MB D2H concurrently per rank from pinned memory, all {self.world} rank(s) at once, "
                                "slowest rank"}
        return self._ceiling

    def pinned(self, arr_T):
        """(rows, T) column-major numpy view of a page-locked copy of arr_T (T, rows)."""
        t = self.torch.from_numpy(np.ascontiguousarray(arr_T)).pin_memory()
        return t, t.numpy().T


def run_fbf(ctx: Ctx, steps: int, warmup: int) -> dict:
    torch, vcb, args = ctx.torch, ctx.vcb, ctx.args
    T = args.frames
    gm, fm = vcb.synth.config_c1(T)                 # same seeded inputs on every rank (weak scaling)
    g = vcb.GMMMap(*gm)
    rows = fm.shape[0]
    dfm = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()   # frame-major (T, rows) == Julia's (rows, T)
    in_bytes = dfm.numel() * 8
    out = [None]

    def step():
        out[0] = vcb.vc(g, dfm)
    # one kernel per step: the step time is the kernel time
    ms, launches, _, _ = ctx.timed(step, steps, warmup)
    value = ctx.world * T / (ms * 1e-3)

    # end to end through the public host API: pinned host buffers, H2D + D2H inside the timed region
    _, hfm = ctx.pinned(fm.T)
    hout_t = torch.empty((T, rows), dtype=torch.float64).pin_memory()
    hout = hout_t.numpy().T
    e2e_steps = max(3, min(steps, 10))
    for _ in range(2):
        vcb.vc(g, hfm, out=hout)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        vcb.vc(g, hfm, out=hout)
    torch.cuda.synchronize()
    e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_ok = bool(np.array_equal(hout[0], fm[0]))
    ceiling = ctx.copy_ceiling()
    if ctx.rank != 0:
        return {}
    tf32_peak = measure_tf32_peak(torch)
    achieved = F_FBF * T / (ms * 1e-3) / 1e12
    used_tc = args.variant != 1
    prof = ncu_summary("prof_fbf_tc" if used_tc else "prof_fbf_simt")
    cfg = config_for("fbf", ctx.world)
    cfg["frames_per_gpu"] = T        # (identical to the --impl reference arm's config unless --frames is given)
    return {
        "metric": METRICS["fbf"], "value": value, "unit": "frames/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3 (fp32-accurate tensor-core split; f64 API)" if used_tc else "f32", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": ctx.world * T / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": in_bytes, "steps": e2e_steps, "power_row_ok": e2e_ok,
                "gbs_each_direction_per_gpu": in_bytes / e2e_s / 1e9, "copy_ceiling": ceiling,
                "fraction_of_copy_ceiling": (in_bytes / e2e_s / 1e9) / ceiling["gbs_each_direction_per_gpu"],
                "note": "vc(g, fm) through the C ABI with pinned Float64 host buffers, pipelined H2D/kernel/D2H; the path "
                        "moves 16(D+1) B per frame each way, so host<->device copy bandwidth is its ceiling"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": achieved / tf32_peak, "traffic": prof["traffic"],
                     "traffic_unit": "bytes/launch (DRAM read+write, ncu)", "traffic_source": prof["source"],
                     "kernel": "gmm_tc_kernel<24,true>" if used_tc else "gmm_simt_kernel<24,2,true>",
                     "kernel_ms": ms, "algorithmic_flop_per_frame": F_FBF,
                     "peak_source": "dense TF32 measured in this run (torch.matmul 8192^3, best of 10); "
                                    f"bf16 {ctx.peaks.get('bf16_tflops')} TF/s, HBM {ctx.peaks.get('hbm_gbs')} GB/s {ctx.peak_src}",
                     "note": "achieved counts ALGORITHMIC flops (4MD^2+2MD per frame); the 3xTF32 split issues 3 MMAs "
                             "per data k-step plus 1 for the offset step (K: 25 -> 80 effective), so the tensor pipe "
                             "executes ~3.3x that; ncu: sm__pipe_tensor_cycles_active "
                             f"{prof.get('sm__pipe_tensor_cycles_active')} %"},
    }


_CACHE: dict = {}


def run_traj(ctx: Ctx, steps: int, warmup: int, which: str = "c4", limit: int = 500) -> dict:
    """which = "c4": the 8192-utterance batch sharded over the ranks (strong scaling, the --path traj
    line); "c2": 1000 utterances x 500 frames with 64 mixtures per GPU (side number)."""
    torch, vcb = ctx.torch, ctx.vcb
    if which == "c4":
        gm = vcb.synth.random_joint_gmm(1004, C4_MIX, 96)
        all_off = np.arange(C4_UTT + 1, dtype=np.int64) * C4_FRAMES
        (ub, ue), off = vcb.shard.shard_ragged(all_off, ctx.rank, ctx.world)
        fm, off = vcb.synth.c4_utterances(gm, np.arange(ub, ue), C4_FRAMES)     # this rank's own utterances
        total_frames = C4_UTT * C4_FRAMES
        cfg = config_for("traj", ctx.world)
        cfg["chunk_limit"] = limit
    else:
        if "c2" not in _CACHE:
            _CACHE["c2"] = vcb.synth.config_c2(1000, 500)
        gm, fm, off = _CACHE["c2"]
        ub, ue = 0, 1000
        total_frames = ctx.world * 1000 * 500
        cfg = {"workload": "C2: 48-dim source (static+delta), 64 mixtures, 1000 utt x 500 frames per GPU, chunk limit "
                           f"{limit}" + (" (the CLI default, bin/vc.jl:18)" if limit == 100 else " (one chunk per utterance)"),
               "chunk_limit": limit}
    local_frames = int(off[-1])
    tj = vcb.TrajectoryGMMMap(vcb.GMMMap(*gm), limit)
    dfm = torch.from_numpy(np.ascontiguousarray(fm.T)).cuda()
    out = [None]

    def step():
        out[0], = vcb.vc_batch(tj, dfm, off, _split=False)
    ms, launches, solver_ms, stages = ctx.timed(step, steps, warmup, stage_index=3)
    vcb.traj_status(tj)
    value = total_frames / (ms * 1e-3)

    gather = None
    if which == "c4" and ctx.world > 1:
        # optional result gather, timed separately from the conversion
        sizes = [int(np.diff(vcb.shard.shard_ragged(all_off, r, ctx.world)[0])[0]) * C4_FRAMES for r in range(ctx.world)]
        for _ in range(2):
            vcb.shard.gather_frames(out[0], ctx.dist, sizes=sizes)
        ctx.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            full = vcb.shard.gather_frames(out[0], ctx.dist, sizes=sizes)
        e.record(); ctx.barrier()
        gms = ctx.max_over_ranks(s.elapsed_time(e) / 5)
        gather = {"ms": gms, "bytes_to_rank0": (total_frames - local_frames) * 25 * 8 if ctx.rank == 0 else None,
                  "gbs_into_rank0": (total_frames - local_frames) * 25 * 8 / (gms * 1e-3) / 1e9 if ctx.rank == 0 else None,
                  "how": "shard.gather_frames: ncclSend/ncclRecv of each rank's (frames, 25) result into rank 0's "
                         "preallocated batch over NVLink/NVSwitch"}
        if ctx.rank == 0:
            gather["rows_ok"] = bool(full is not None and full.shape[0] == total_frames)
        del full

    # end to end: host Float64 buffers through vcb_traj_vc_batch (H2D + kernels + D2H, sliced pipeline)
    _, hfm = ctx.pinned(fm.T)
    hout_t = torch.empty((local_frames, 25), dtype=torch.float64).pin_memory()
    hout = hout_t.numpy().T
    e2e_steps = 3
    for _ in range(2):
        vcb.vc_batch(tj, hfm, off, _split=False, out=hout)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        vcb.vc_batch(tj, hfm, off, _split=False, out=hout)
    torch.cuda.synchronize()
    e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    h2d = ctx.sum_over_ranks(float(hfm.size * 8)) / ctx.world
    d2h = ctx.sum_over_ranks(float(hout.size * 8)) / ctx.world
    dev = out[0].cpu().numpy().T
    same = bool(np.array_equal(dev, hout)) and bool(np.array_equal(hout[0], fm[0]))
    if ctx.rank != 0:
        return {}
    hbm = ctx.peaks.get("hbm_gbs", 6650.0)
    prof = ncu_summary("prof_traj")
    kms = solver_ms if solver_ms else ms
    achieved = B_TRAJ * local_frames / (kms * 1e-3) / 1e9
    return {
        "metric": METRICS["traj"] if which == "c4" else "converted frames/sec (trajectory, C2)",
        "value": value, "unit": "frames/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if which == "c4" else "weak", "vs_baseline": None,
        "dtype": "f64 (band solve, E/PE) + tf32x3 arg-max with f64 re-check", "data": "synthetic",
        "config": cfg,
        "shard": {"rank0_utterances": [int(ub), int(ue)], "rank0_frames": local_frames},
        "stages_ms": {"argmax_recheck": stages[0] if len(stages) > 0 else None, "bucketing": stages[1] if len(stages) > 1 else None,
                      "e_pe_panels": stages[2] if len(stages) > 2 else None, "band_solver": solver_ms},
        "gather": gather,
        "e2e": {"value": total_frames / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "matches_device_path": same,
                "note": "vc(c, fms) batch through vcb_traj_vc_batch with pinned Float64 host buffers; utterance slices of "
                        "one solver wave rotate through streams (bytes are the per-rank average)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": (prof["traffic"] / (prof["launch__grid_size"] * 500.0) * local_frames
                                 if prof["traffic"] and prof.get("launch__grid_size") else None),
                     "traffic_unit": "bytes/launch (DRAM read+write; ncu capture of a 500-frame-per-chunk launch, scaled by "
                                     "frames to this launch)",
                     "traffic_source": prof["source"], "kernel": "traj_solve_warp<3,true>", "kernel_ms": kms,
                     "algorithmic_bytes_per_frame": B_TRAJ, "frames_per_launch": local_frames,
                     "step_frac": B_TRAJ * local_frames / (ms * 1e-3) / 1e9 / hbm,
                     "peak_source": f"HBM copy bandwidth {ctx.peak_src}",
                     "fp64_pipe": {"pipe_cycles_per_frame": 240 * 16.1 + 900 * 2.2,
                                   "floor_ms": (240 * 16.1 + 900 * 2.2) * local_frames / (4 * 148) / ((ctx.peaks.get("sm_max_mhz") or 1965.0) * 1e3),
                                   "frac_of_floor": (240 * 16.1 + 900 * 2.2) * local_frames / (4 * 148) / ((ctx.peaks.get("sm_max_mhz") or 1965.0) * 1e3) / kms,
                                   "ncu_pipe_active_pct": {"dmma": prof.get("sm__pipe_tensor_cycles_active"), "fp64": prof.get("sm__pipe_fp64_cycles_active")},
                                   "note": "the real bound: DMMA (16.1 cycles per sub-partition) and DFMA (2.2) share the FP64 pipe "
                                           "(tools/micro/dmma_bench.cu); ~240 DMMA + ~900 FP64 instructions per frame, 592 sub-partitions"},
                     "note": "achieved = 28,224 B/frame (SURVEY 8d: L written + read once, E read, y written) x frames of "
                             "one launch / the band solver's own duration (CUDA events around the kernel inside the timed "
                             "region); step_frac divides by the whole step (arg-max + E/PE + solver) instead. The solver "
                             "stores only Linv_t and L[t][t-1] (15,360 B/frame), so its real DRAM traffic is below the "
                             f"algorithmic figure; ncu: warps active {prof.get('sm__warps_active')} %, "
                             f"FP64/DMMA pipe {prof.get('sm__pipe_fp64_cycles_active')} %"},
    }


def run_dtw(ctx: Ctx, steps: int, warmup: int) -> dict:
    torch, vcb = ctx.torch, ctx.vcb
    tm, to, sq, so = dtw_pairs_for_rank(vcb, ctx.rank)
    dtm = torch.from_numpy(np.ascontiguousarray(tm.T)).cuda()
    dsq = torch.from_numpy(np.ascontiguousarray(sq.T)).cuda()
    d = vcb.DTWs.DTW(fstep=0, bstep=2)
    cells = float(np.sum(np.diff(to).astype(np.float64) * np.diff(so)))
    res = [None]

    def step():
        res[0] = vcb.DTWs.fit_batch(d, dtm, to, dsq, so)
    ms, launches, kernel_ms, _ = ctx.timed(step, steps, warmup, stage_index=1)
    total_cells = ctx.sum_over_ranks(cells)
    value = total_cells / (ms * 1e-3)

    # end to end: host arrays through vcb_dtw_fit_batch
    _, htm = ctx.pinned(tm.T)
    _, hsq = ctx.pinned(sq.T)
    for _ in range(3):          # the first calls grow the stream-ordered pool and the page-locked bounce buffer
        vcb.DTWs.fit_batch(d, htm, to, hsq, so)
    ctx.barrier()
    e2e_steps = 10
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hp, hc = vcb.DTWs.fit_batch(d, htm, to, hsq, so)
    e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    same = bool(np.array_equal(res[0][0].cpu().numpy(), hp))
    if ctx.rank != 0:
        return {}
    hbm = ctx.peaks.get("hbm_gbs", 6650.0)
    prof = ncu_summary("prof_dtw")
    kms = kernel_ms if kernel_ms else ms
    achieved = B_DTW * cells / (kms * 1e-3) / 1e9
    sm_clock = (ctx.peaks.get("sm_max_mhz") or 1965.0) * 1e6
    fp64_floor_ms = DP_OPS_PER_CELL * cells / (64.0 * 148 * sm_clock) * 1e3
    return {
        "metric": METRICS["dtw"], "value": value, "unit": "cells/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (bit-exact)",
        "data": "synthetic", "config": config_for("dtw", ctx.world),
        "e2e": {"value": total_cells / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": int((tm.size + sq.size) * 8),
                "d2h_bytes_per_step": int(hp.size * 8 + hc.size * 8), "steps": e2e_steps, "matches_device_path": same,
                "note": "DTWs.fit_batch through vcb_dtw_fit_batch with pinned Float64 host buffers"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": prof["traffic"], "traffic_unit": "bytes/launch (DRAM read+write, ncu)",
                     "traffic_source": prof["source"], "kernel": "dtw_stream_kernel", "kernel_ms": kms,
                     "algorithmic_bytes_per_cell": B_DTW, "cells_per_launch": cells,
                     "peak_source": f"HBM copy bandwidth {ctx.peak_src}",
                     "fp64_pipe": {"dp_ops_per_cell": DP_OPS_PER_CELL, "floor_ms": fp64_floor_ms, "frac_of_floor": fp64_floor_ms / kms,
                                   "ncu_pipe_fp64_active_pct": prof.get("sm__pipe_fp64_cycles_active"),
                                   "note": "64 FP64 lanes/clk/SM x 148 SMs at the maximum SM clock"},
                     "note": "achieved uses SURVEY 8d's 17 B/cell of the two-pass form (cost matrix written + read, 1 B "
                             "back-pointer); the kernel is fused (persistent warp pipeline, cost matrix never leaves the "
                             "SM), so the real bound is the FP64 pipe: see fp64_pipe"},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="fbf", choices=["fbf", "traj", "dtw"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU per step (fbf; default: the C1 size)")
    ap.add_argument("--skip-extras", action="store_true", help="fbf: skip the other paths' sub-lines and the CPU baseline")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 CUDA-core kernel, 2 tcgen05 kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        cpu_reference_arm(args)
        return

    ctx = Ctx(args)
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()       # samples cover warm-up, the timed device loop and the end-to-end loop
    if args.path == "fbf":
        line = run_fbf(ctx, args.steps, args.warmup)
    elif args.path == "traj":
        line = run_traj(ctx, args.steps, args.warmup, "c4", C4_FRAMES)
    else:
        line = run_dtw(ctx, args.steps, args.warmup)
    clocks = sampler.stop() if ctx.rank == 0 else None
    if ctx.rank == 0:
        line["clocks"] = clocks
        line["host"] = ctx.host

    # the other paths' complete sub-lines (same structure as a contract line), default run only
    if args.path == "fbf" and not args.skip_extras:
        extras = {}
        side_steps = max(3, min(args.steps, 5))
        for name, fn in (("trajectory_c4", lambda: run_traj(ctx, side_steps, 3, "c4", C4_FRAMES)),
                         ("trajectory_c2", lambda: run_traj(ctx, side_steps, 3, "c2", 500)),
                         ("trajectory_c2_limit100", lambda: run_traj(ctx, side_steps, 3, "c2", 100)),
                         ("dtw_c3", lambda: run_dtw(ctx, side_steps, 3))):
            try:
                extras[name] = fn()
            except Exception as ex:  # side numbers must never break the contract line
                extras[name] = {"error": repr(ex)}
                ctx.torch.cuda.synchronize()
        if ctx.rank == 0:
            if ctx.world == 1:
                for name, path in (("trajectory_c4", "traj"), ("dtw_c3", "dtw")):
                    if "error" not in extras[name]:
                        extras[name]["cpu_baseline"] = cpu_baseline(path)
            line["other_paths"] = extras
    if ctx.rank == 0:
        if ctx.world == 1 and not (args.path == "fbf" and args.skip_extras):
            line["cpu_baseline"] = cpu_baseline(args.path)
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNITS[args.path], "cores": 0, "kind": "port",
                                    "sample": "not measured in this run (N>1 or --skip-extras)"}
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
