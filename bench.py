#!/usr/bin/env python
"""bench.py -- fastMPC solves/sec (fp64, batched) on N B200s, next to the host-CPU port.

A "step" = one batched `mpc_fixed_log_newton(niters=5, kappa=0.01)` over 4096 independent
VAR(2) controller instances per GPU (n = 28 modes, m = 144 actuators, horizon T = 20;
BASELINE.json configs[1], README regime: Q = 1.5e4 I, R = I, |u| <= 28, warm starts).
Instances are independent: they shard across ranks with NO data-path collective (weak scaling,
4096 instances per GPU); NCCL only carries the timing/statistics reduction.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                          (CPU port on the host cores)

One JSON line on rank 0; see DESIGN.md "Measurement" for every key.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ZERN, T_HOR, NB_PER_GPU, NITERS, KAPPA = 6, 20, 4096, 5, 0.01
NSETS = 3           # rotating input sets: 3 x 150 MB in + 113 MB out per step >> 126 MB L2


def f_newton(n, m, T):
    """Algorithmic flops per Newton iteration per instance."""
    return T * (n * n * m + (19.0 / 3.0) * n ** 3 + 8 * n * m + 26 * n * n)


def workload_desc():
    return {"workload": "VAR(2) fastMPC, n=28 modes (N=6), m=144, T=20, 4096 instances per GPU, niters=5, kappa=0.01, "
                        "README weights/bounds, warm starts (BASELINE.json configs[1])",
            "n": 28, "m": 144, "T": T_HOR, "instances_per_gpu": NB_PER_GPU, "niters": NITERS, "kappa": KAPPA,
            "sharding": "instances split across ranks, no data-path collective",
            "l2_policy": f"inputs larger than L2: {NSETS} rotating input sets, ~263 MB touched per step"}


def make_inputs(p, nb, rank):
    from mpc_sensorlessao_b200 import synth
    return [synth.warm_inputs(p, nb, seed=100 + 10 * rank + s) for s in range(NSETS)]


def cpu_port_rate(p, sets, nsample, nthreads=0, reps=1):
    """Times the structured C oracle (oracle/fmpc_ref.c, OpenMP over instances) on `nsample` instances."""
    from oracle import fmpc_ref
    wi = sets[0]
    nb = min(nsample, wi["x0"].shape[0])
    z0 = np.concatenate([wi["U0"][:nb], wi["X0"][:nb]], axis=2).reshape(nb, -1)
    args = (p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, KAPPA, NITERS, wi["x0"][:nb].T, wi["x0_pre"][:nb].T, None,
            z0.T, wi["nu0"][:nb].T)
    fmpc_ref.solve_batch(*args, nthreads=nthreads)          # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fmpc_ref.solve_batch(*args, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / reps
    return nb / dt, int(out["iters"].sum()), nb, dt


try:
    ORIG_AFFINITY = os.sched_getaffinity(0)      # before any NUMA binding of this rank
except AttributeError:
    ORIG_AFFINITY = None


def host_threads():
    return len(ORIG_AFFINITY) if ORIG_AFFINITY else (os.cpu_count() or 1)


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e path are
    allocated on the GPU's own NUMA node (with several ranks per host the copies otherwise cross the socket interconnect).
    Returns the number of CPUs bound to, or None if NVML / affinity is unavailable."""
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        h = None
        for i in range(nv.nvmlDeviceGetCount()):
            hh = nv.nvmlDeviceGetHandleByIndex(i)
            if nv.nvmlDeviceGetPciInfo(hh).bus == bus:
                h = hh
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): an in-process NVML thread
    (nvidia_ml_py, one sample every ~5 ms -- the timed region is only ~60 ms long), nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.sm, self.smax, self.reasons, self.stop_flag, self.thread, self.how = [], [], set(), False, None, None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve the NVML handle through the PCI bus id of the CUDA device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hh = nv.nvmlDeviceGetHandleByIndex(i)
                    if nv.nvmlDeviceGetPciInfo(hh).bus == bus:
                        h = hh
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = [float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))]
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            if self.sm:
                return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.smax)), "reasons": sorted(self.reasons),
                        "samples": len(self.sm), "how": "nvml, 5 ms period, inside the timed region"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm),
                "how": "nvidia-smi -lms 50"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is MATLAB
    (no MATLAB/Octave on the box, no C core to compile -- SURVEY.md F1/F2), so this arm times the
    structured C/OpenMP restatement (oracle/fmpc_ref.c, kind = "port") with every host thread."""
    if rank != 0:
        return
    import mpc_sensorlessao_b200  # noqa: F401
    from mpc_sensorlessao_b200 import synth
    from oracle import fmpc_ref
    p = synth.make_problem(N_ZERN, T_HOR)
    nsample = 2048
    sets = [synth.warm_inputs(p, nsample, seed=100)]
    cores = host_threads()          # torchrun exports OMP_NUM_THREADS=1: ask for every host thread explicitly
    if args.warmup > 0:
        cpu_port_rate(p, sets, nsample, nthreads=cores, reps=args.warmup)                      # W untimed warm-up steps
    rate, it1, nb, dt1 = cpu_port_rate(p, sets, nsample, nthreads=cores, reps=max(args.steps, 1))   # K timed steps (mean)
    line = {"impl": "reference", "metric": "fastmpc_solves_per_sec", "value": rate, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt1 * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_desc(),
            "cpu_baseline": {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{nb} instances of the workload per step, OpenMP over instances, all {cores} host threads; "
                                       "structured C restatement of Fast_MPC/VAR_2 (the MATLAB reference cannot run here)"},
            "e2e": {"value": rate, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "newton_iters_per_solve": it1 / nb}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=NB_PER_GPU, help="instances per GPU (default: the metric's 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import mpc_sensorlessao_b200 as pk
    from mpc_sensorlessao_b200 import synth
    from mpc_sensorlessao_b200._lib import load_library

    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (NCCL's version banner, warnings) is sent to
    # stderr at the file-descriptor level and the real stdout is restored just before the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch decides where the pages live
    if world > 1:
        # NCCL prints its version banner to STDOUT when NCCL_DEBUG is VERSION/WARN: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)
    L = load_library()
    nb = args.instances
    p = synth.make_problem(N_ZERN, T_HOR)
    n, m, T = p.n, p.m, p.T
    sets = make_inputs(p, nb, rank)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T, p.x_min, p.x_max, max_batch=nb, device=local_rank)
    params = hb.params(KAPPA, NITERS, 0)

    # FP64 pipe peaks, measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)
    peak_dmma = pk.fp64_peak(local_rank, 1, 4000)
    peak_dfma = pk.fp64_peak(local_rank, 0, 4000)
    peak = max(peak_dmma, peak_dfma)

    # ---- device-resident inputs (value) and pinned host inputs (e2e) ----
    keys = ("x0", "x0_pre", "X0", "U0", "nu0")
    dsets = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).to(dev) for k in keys} for s in sets]
    hsets = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).pin_memory() for k in keys} for s in sets]
    dX = torch.empty((nb, T, n), dtype=torch.float64, device=dev)
    dU = torch.empty((nb, T, m), dtype=torch.float64, device=dev)
    dstat = torch.empty(nb, dtype=torch.int32, device=dev)
    dit = torch.empty(nb, dtype=torch.int32, device=dev)
    hX = torch.empty((nb, T, n), dtype=torch.float64).pin_memory()
    hU = torch.empty((nb, T, m), dtype=torch.float64).pin_memory()
    hstat = torch.empty(nb, dtype=torch.int32).pin_memory()
    hit = torch.empty(nb, dtype=torch.int32).pin_memory()
    # a dedicated (non-default) stream: the library treats stream 0/NULL as "use the handle's own stream",
    # and the per-launch CUDA events below must sit on the stream the kernel is launched on
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def step_dev(i):
        d = dsets[i % NSETS]
        rc = L.fmpc_step_d(hb._h, C.byref(params), nb, vp(d["x0"]), vp(d["x0_pre"]), None, None, None, vp(d["X0"]), vp(d["U0"]),
                           vp(d["nu0"]), vp(dX), vp(dU), vp(dstat), vp(dit), C.c_void_p(stream.cuda_stream))
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))

    def step_host(i):
        d = hsets[i % NSETS]
        rc = L.fmpc_step(hb._h, C.byref(params), nb, vp(d["x0"]), vp(d["x0_pre"]), None, None, None, vp(d["X0"]), vp(d["U0"]),
                         vp(d["nu0"]), vp(hX), vp(hU), vp(hstat), vp(hit), None)
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up; Newton iterations per input set (constant across repeats) ----
    iters_per_set = []
    for i in range(max(args.warmup, NSETS)):
        step_dev(i)
        torch.cuda.synchronize(dev)
        if i < NSETS:
            iters_per_set.append(hb.last_newton_iters())
    status_hist = np.bincount(dstat.cpu().numpy(), minlength=5).tolist()

    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled meanwhile ----
    K = args.steps
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    sampler = ClockSampler(local_rank)
    launches0 = hb.launch_count
    barrier()
    sampler.start()
    evs[0].record(stream)
    for i in range(K):
        step_dev(i)
        evs[i + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = hb.launch_count - launches0
    step_ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
    total_ms = evs[0].elapsed_time(evs[K])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    newton_iters = sum(iters_per_set[i % NSETS] for i in range(K))

    # ---- e2e: same metric through the host-buffer C-ABI call (H2D + solve + D2H inside the timed region) ----
    for i in range(2):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s_max = float(te.item())
    h2d = sum(hsets[0][k].numel() * 8 for k in keys)
    d2h = hX.numel() * 8 + hU.numel() * 8 + hstat.numel() * 4 + hit.numel() * 4

    # ---- gather statistics (NCCL carries only this) ----
    stats = torch.tensor([float(newton_iters), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    newton_iters_all, launches_all = float(stats[0].item()), int(stats[1].item())

    if rank == 0:
        F = f_newton(n, m, T)
        kname = {2: "fmpc_solve_kernel_warp<28,4>", 1: "fmpc_solve_kernel_mma<28>", 0: "fmpc_solve_kernel_v1"}.get(hb.kernel_kind, "?")
        traffic = None      # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if kname in tj and nb == NB_PER_GPU:
                traffic = float(tj[kname]["dram_bytes_per_launch"])
        except Exception:
            traffic = None
        solves = nb * K * world
        value = solves / (total_ms_max * 1e-3)
        # roofline of the (single) solve kernel: algorithmic flops of one launch / its average duration
        flops_per_launch = (newton_iters / K) * F
        kern_ms = float(step_ms.mean())
        achieved = flops_per_launch / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": "fastmpc_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_desc(),
            "clocks": clocks,
            "e2e": {"value": solves / e2e_s_max, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "fmpc_step (C-ABI, pinned host buffers)", "numa_bound_cpus": numa},
            "gpu_launches": launches_all,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA)",
                         "kernel": kname, "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                         "peak_source": "measured live: fmpc_fp64_peak DMMA m8n8k4 %.2f / DFMA %.2f TFLOP/s "
                                        "(MEASURED_PEAKS.json has no FP64 entry)" % (peak_dmma, peak_dfma),
                         "flops_per_newton_iter": F, "newton_iters_per_launch": newton_iters / K, "kernel_ms": kern_ms},
            "newton_iters_per_solve": newton_iters_all / solves, "status_hist": status_hist,
            "aggregate_tflops": newton_iters_all * F / (total_ms_max * 1e-3) / 1e12,
        }
        if world == 1 and not args.no_cpu_baseline:
            if ORIG_AFFINITY:
                os.sched_setaffinity(0, ORIG_AFFINITY)      # the CPU baseline uses every host thread again
            cores = host_threads()
            nsample = nb
            r0, _, _, dt0 = cpu_port_rate(p, sets, nsample, nthreads=cores, reps=1)
            reps = max(3, min(60, int(round(12.0 / max(dt0, 1e-3)))))          # ~12 s of CPU work
            rate, it, nbs, dt = cpu_port_rate(p, sets, nsample, nthreads=cores, reps=reps)
            line["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                                    "sample": f"all {nbs} instances of one step, {reps} repetitions ({reps * dt:.1f} s), structured "
                                              "C/OpenMP port (oracle/fmpc_ref.c) on every host thread; the MATLAB reference cannot "
                                              "run on this box"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    hb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
