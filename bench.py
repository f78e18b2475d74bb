#!/usr/bin/env python
"""bench.py -- fastMPC solves/sec (fp64, batched) on N B200s, next to the host-CPU port.

A "step" = one control step of 4096 independent VAR(2) closed loops per GPU (BASELINE.json configs[1]: n = 28 modes,
m = 144 actuators, horizon T = 20; README regime Q = 1.5e4 I, R = I, |u| <= 28): one batched
`mpc_fixed_log_newton(niters=5, kappa=0.01)` warm-started from the previous solution shifted one stage, the solver state
resident on the device (fmpc_step_r): x0 goes in, the applied input U(:,0) comes out.
  value     the loop with x0 / U(:,0) in HBM (fmpc_step_r_d), CUDA events around every step
  e2e       the same loop through fmpc_step_r with HOST buffers (pinned); e2e_pageable: ordinary malloc'd buffers
  e2e_full  the full-surface call fmpc_step (whole warm start in, whole horizon out), pinned and pageable
Instances are independent: they shard across ranks with NO data-path collective; NCCL only carries the timing /
statistics reduction.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c2_tight|c4|c5|zmf|closed_loop]   (N > 1: under torchrun)
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

KAPPA, NITERS = 0.01, 5
NSETS = 3           # rotating input sets of the full-surface legs: 3 x 150 MB in + 113 MB out per step >> 126 MB L2
PRE = 4             # untimed closed-loop steps in front of the warm-up (cold start at step 0, warm starts settle)

# name -> (Zernike order N, horizon T, instances per GPU (None: 65536 / n_gpus), |u| bound, scaling, what)
CONFIGS = {
    "c2": (6, 20, 4096, 28.0, "weak", "BASELINE.json configs[1]"),
    "c2_tight": (6, 20, 4096, 1.0, "weak", "configs[1] with |u| <= 1: active barrier, several Newton steps per solve"),
    "c4": (6, 20, None, 28.0, "strong", "BASELINE.json configs[3]: 65536 instances sharded across the GPUs"),
    "c5": (10, 30, 16384, 28.0, "weak", "BASELINE.json configs[4]"),
}


def f_newton(n, m, T):
    """Algorithmic flops per Newton iteration per instance."""
    return T * (n * n * m + (19.0 / 3.0) * n ** 3 + 8 * n * m + 26 * n * n)


def workload_desc(name, p, nb, world, mode):
    N, T, _, ub, scaling, what = CONFIGS[name]
    loop = ("closed loops in steady state, plant = model + process noise (north-star d): per step x0 in, U(:,0) out, "
            "warm start = previous solution shifted, kept on the device")
    sets = "independent warm-started solves on rotating input sets (full warm start in, full horizon out)"
    return {"workload": f"VAR(2) fastMPC, n={p.n} modes (N={N}), m={p.m}, T={p.T}, {nb} instances per GPU, niters={NITERS}, "
                        f"kappa={KAPPA}, README weights, |u| <= {ub:g}; {loop if mode == 'loop' else sets} ({what})",
            "name": name, "n": p.n, "m": p.m, "T": p.T, "instances_per_gpu": nb, "instances_total": nb * world, "niters": NITERS,
            "kappa": KAPPA, "u_bound": ub, "mode": mode,
            "sharding": "instances split across ranks, no data-path collective",
            "l2_policy": ("working set larger than L2: every step reads and rewrites the resident horizons (X, U) of all instances "
                          f"({nb * p.T * (p.n + p.m) * 8 / 1e6:.0f} MB) and streams the solver's factor scratch (see workspace_mb)"
                          if mode == "loop" else f"inputs larger than L2: {NSETS} rotating input sets")}


def make_inputs(p, nb, rank, nsets=NSETS):
    from mpc_sensorlessao_b200 import synth
    return [synth.warm_inputs(p, nb, seed=100 + 10 * rank + s) for s in range(nsets)]


def loop_noise(p, nb, K, seed):
    """Process noise of the closed loops: the innovations of synth.aberrations' VAR(2) sequences, (nb, K, n)."""
    from mpc_sensorlessao_b200 import synth
    a = synth.aberrations(p, nb, K + 2, seed=seed)
    return a[:, 2:] - a[:, 1:-1] @ p.A1_true.T - a[:, :-2] @ p.A2_true.T


def cpu_port_rate(p, sets, nsample, nthreads=0, reps=1):
    """Times the structured C oracle (oracle/fmpc_ref.c, OpenMP over instances) on `nsample` instances of a warm input set."""
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


def cpu_loop_rate(p, nb, steps, warm, nthreads, seed=100):
    """The same closed-loop workload on the host: the structured C oracle solves every step (timed), numpy does the plant
    update and the one-stage shift of the warm start (not timed).  Returns (solves/s, newton iters, solves, seconds/step)."""
    from oracle import fmpc_ref
    n, m, T = p.n, p.m, p.T
    noise = loop_noise(p, nb, PRE + warm + steps, seed)
    rs = np.random.RandomState(seed + 1)
    x = noise[:, 0].copy(); xp = np.zeros((nb, n)); z = None
    tsum, its, cnt = 0.0, 0, 0
    for k in range(PRE + warm + steps):
        nu = rs.random_sample((nb, T * n))
        if z is None:
            stage = np.concatenate([(p.u_min + p.u_max) / 2, (p.x_min + p.x_max) / 2])
            z0 = np.tile(stage, T)[None].repeat(nb, 0)
        else:
            Z = z.reshape(nb, T, n + m)
            z0 = np.concatenate([Z[:, 1:], Z[:, -1:]], axis=1).reshape(nb, -1)
        t0 = time.perf_counter()
        out = fmpc_ref.solve_batch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, KAPPA, NITERS, x.T, xp.T, None, z0.T, nu.T,
                                   nthreads=nthreads)
        dt = time.perf_counter() - t0
        z = out["z"].T.copy()
        u = z[:, :m]
        xn = x @ p.A1.T + xp @ p.A2.T + u @ p.B.T + noise[:, min(k + 1, noise.shape[1] - 1)]
        xp, x = x, xn
        if k >= PRE + warm:
            tsum += dt; its += int(out["iters"].sum()); cnt += nb
    return cnt / tsum, its, cnt, tsum / max(steps, 1)


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
    structured C/OpenMP restatement (oracle/fmpc_ref.c, kind = "port") with every host thread, on the same
    closed-loop workload as the GPU arm (a bounded sample of its instances)."""
    if rank != 0:
        return
    import mpc_sensorlessao_b200  # noqa: F401
    from mpc_sensorlessao_b200 import synth
    name = args.config if args.config in CONFIGS else "c2"
    N, T, nbc, ub, scaling, _ = CONFIGS[name]
    nb_gpu = args.instances or nbc or 65536 // max(args.gpus, 1)
    p = synth.make_problem(N, T, u_bound=ub)
    cores = host_threads()          # torchrun exports OMP_NUM_THREADS=1: ask for every host thread explicitly
    nsample = min(nb_gpu, 2048 if p.n <= 32 else 256)
    rate, its, cnt, dt1 = cpu_loop_rate(p, nsample, max(args.steps, 1), max(args.warmup, 0), cores)
    line = {"impl": "reference", "metric": "fastmpc_solves_per_sec", "value": rate, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt1 * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_desc(name, p, nb_gpu, args.gpus, "loop"),
            "cpu_baseline": {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{nsample} of the workload's closed loops per step, OpenMP over instances, all {cores} host "
                                       "threads; structured C restatement of Fast_MPC/VAR_2 (the MATLAB reference cannot run here); "
                                       "solve time only (plant update and warm-start shift in numpy are not timed)"},
            "e2e": {"value": rate, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "newton_iters_per_solve": its / max(cnt, 1)}
    print(json.dumps(line), flush=True)


class Ctx:
    """Per-process CUDA / distributed state shared by the measurement legs."""
    pass


def sync_max(cx, seconds):
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(cx):
    import torch
    import torch.distributed as dist
    if cx.world > 1:
        dist.barrier()
    torch.cuda.synchronize(cx.dev)


def kernel_name(hb, n):
    npot = 8 if n <= 8 else 16 if n <= 16 else 24 if n <= 24 else 28 if n <= 28 else 32
    return {2: f"fmpc_solve_kernel_warp<{npot},{n // 8 + 1}>", 1: f"fmpc_solve_kernel_mma<{(n + 7) // 8 * 8}>",
            0: "fmpc_solve_kernel_v1", 3: "fmpc_solve_kernel_gen"}.get(hb.kernel_kind, "?")


def traffic_of(kname, nb, mode):
    """DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = tj.get(kname)
        if e and int(e.get("instances", 4096)) == nb and e.get("mode", "sets") == mode:
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def measure_loop(cx, name, p, nb, K, W, hostlegs=True):
    """Closed loops with the solver state resident on the device.  Returns the measurements of the `value` leg and, if
    `hostlegs`, of the e2e legs through fmpc_step_r with pinned and pageable host buffers."""
    import torch
    import mpc_sensorlessao_b200 as pk
    L, dev = cx.L, cx.dev
    n, m, T = p.n, p.m, p.T
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T, p.x_min, p.x_max, max_batch=nb, device=cx.local_rank)
    params = hb.params(KAPPA, NITERS, 0)
    KT = PRE + W + K
    noise = loop_noise(p, nb, KT + 1, seed=100 + 10 * cx.rank)
    # pass 0 (untimed): run the loops once, the plant on the host, to record the x0 sequence every leg replays
    x0s = np.empty((KT, nb, n))
    x = noise[:, 0].copy(); xp = np.zeros((nb, n))
    for k in range(KT):
        x0s[k] = x
        out = hb.step_resident(x, reset=(k == 0), params=params)
        xn = x @ p.A1.T + xp @ p.A2.T + out["u0"] @ p.B.T + noise[:, k + 1]
        xp, x = x, xn
    vp = lambda t: C.c_void_p(t.data_ptr())
    res = {"kernel": kernel_name(hb, n), "rms_x0": float(np.sqrt((x0s[PRE + W:] ** 2).mean())), "u0_max": float(np.abs(out["u0"]).max()),
           "ws_mb": hb.workspace_bytes / 1e6}

    # ---- value: x0 sequence and outputs in HBM, one CUDA event pair per step on the launching stream ----
    dx0 = torch.from_numpy(x0s).to(dev)
    du0 = torch.empty((nb, m), dtype=torch.float64, device=dev)
    dst = torch.empty((K + 1, nb), dtype=torch.int32, device=dev)
    dit = torch.zeros((K + 1, nb), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)

    # dual starts of the value leg: resident in HBM like every other input (NSETS rotating arrays); the host legs below use
    # nu0 = NULL instead, i.e. the handle's MATLAB stream generated on the device (nothing to ship)
    dnu = torch.rand((NSETS, nb, T * n), dtype=torch.float64, device=dev)

    def step_dev(k, slot):
        rc = L.fmpc_step_r_d(hb._h, C.byref(params), nb, 1 if k == 0 else 0, vp(dx0[k]), None, None, None, None, vp(dnu[k % NSETS]),
                             vp(du0), None, None, vp(dst[slot]), vp(dit[slot]), C.c_void_p(stream.cuda_stream))
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))

    for k in range(PRE + W):
        step_dev(k, K)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    sampler = ClockSampler(cx.local_rank)
    launches0 = hb.launch_count
    barrier(cx)
    sampler.start()
    evs[0].record(stream)
    for i in range(K):
        step_dev(PRE + W + i, i)
        evs[i + 1].record(stream)
    barrier(cx)
    res["clocks"] = sampler.stop()
    res["launches"] = hb.launch_count - launches0
    step_ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
    res["kern_ms"] = float(step_ms.mean())
    res["total_ms"] = sync_max(cx, evs[0].elapsed_time(evs[K]))
    res["newton_iters"] = int(dit[:K].sum().item())
    res["iters_hist"] = np.bincount(dit[:K].cpu().numpy().reshape(-1), minlength=NITERS + 1).tolist()
    res["status_hist"] = np.bincount(dst[:K].cpu().numpy().reshape(-1), minlength=5).tolist()
    res["nb"], res["K"] = nb, K

    # ---- e2e: the same loop through the host-buffer call; x0 from host memory in, U(:,0) + status + iters out, every step ----
    def host_leg(pinned):
        if pinned:
            hx = [torch.from_numpy(x0s[k].copy()).pin_memory() for k in range(KT)]
            hu = torch.empty((nb, m), dtype=torch.float64).pin_memory()
            hs, hi = torch.empty(nb, dtype=torch.int32).pin_memory(), torch.empty(nb, dtype=torch.int32).pin_memory()
            ptr = lambda t: C.c_void_p(t.data_ptr())
        else:
            hx = [x0s[k].copy() for k in range(KT)]
            hu, hs, hi = np.empty((nb, m)), np.empty(nb, dtype=np.int32), np.empty(nb, dtype=np.int32)
            ptr = lambda t: C.c_void_p(t.ctypes.data)

        def step_host(k):
            rc = L.fmpc_step_r(hb._h, C.byref(params), nb, 1 if k == 0 else 0, ptr(hx[k]), None, None, None, None, None, ptr(hu),
                               None, None, ptr(hs), ptr(hi), None)
            if rc:
                raise pk.FmpcError(rc, pk.strerror(rc))
        for k in range(PRE + W):
            step_host(k)
        barrier(cx)
        t0 = time.perf_counter()
        for i in range(K):
            step_host(PRE + W + i)
        torch.cuda.synchronize(dev)
        return sync_max(cx, time.perf_counter() - t0)

    if hostlegs:
        res["e2e_s"] = host_leg(True)
        res["e2e_pageable_s"] = host_leg(False)
        res["h2d"] = nb * n * 8
        res["d2h"] = nb * m * 8 + nb * 4 + nb * 4
    hb.close()
    return res


def measure_sets(cx, name, p, nb, K, W, nsets=NSETS, hostlegs=True):
    """Independent warm-started solves on rotating input sets: fmpc_step_d with everything in HBM, and the full-surface
    host call fmpc_step (whole warm start in, whole horizon out) with pinned and with pageable buffers."""
    import torch
    import mpc_sensorlessao_b200 as pk
    L, dev = cx.L, cx.dev
    n, m, T = p.n, p.m, p.T
    sets = make_inputs(p, nb, cx.rank, nsets)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T, p.x_min, p.x_max, max_batch=nb, device=cx.local_rank)
    params = hb.params(KAPPA, NITERS, 0)
    keys = ("x0", "x0_pre", "X0", "U0", "nu0")
    dsets = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).to(dev) for k in keys} for s in sets]
    dX = torch.empty((nb, T, n), dtype=torch.float64, device=dev)
    dU = torch.empty((nb, T, m), dtype=torch.float64, device=dev)
    dstat = torch.empty(nb, dtype=torch.int32, device=dev)
    dit = torch.empty(nb, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def step_dev(i):
        d = dsets[i % nsets]
        rc = L.fmpc_step_d(hb._h, C.byref(params), nb, vp(d["x0"]), vp(d["x0_pre"]), None, None, None, vp(d["X0"]), vp(d["U0"]),
                           vp(d["nu0"]), vp(dX), vp(dU), vp(dstat), vp(dit), C.c_void_p(stream.cuda_stream))
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))

    iters_per_set, hist = [], np.zeros(NITERS + 1, dtype=np.int64)
    for i in range(max(W, nsets)):
        step_dev(i)
        torch.cuda.synchronize(dev)
        if i < nsets:
            iters_per_set.append(hb.last_newton_iters())
            hist += np.bincount(dit.cpu().numpy(), minlength=NITERS + 1)[:NITERS + 1]
    res = {"kernel": kernel_name(hb, n), "status_hist": np.bincount(dstat.cpu().numpy(), minlength=5).tolist(),
           "iters_hist": hist.tolist()}
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    sampler = ClockSampler(cx.local_rank)
    launches0 = hb.launch_count
    barrier(cx)
    sampler.start()
    evs[0].record(stream)
    for i in range(K):
        step_dev(i)
        evs[i + 1].record(stream)
    barrier(cx)
    res["clocks"] = sampler.stop()
    res["launches"] = hb.launch_count - launches0
    res["kern_ms"] = float(np.mean([evs[i].elapsed_time(evs[i + 1]) for i in range(K)]))
    res["total_ms"] = sync_max(cx, evs[0].elapsed_time(evs[K]))
    res["newton_iters"] = sum(iters_per_set[i % nsets] for i in range(K))
    res["nb"], res["K"], res["sets"] = nb, K, sets

    def host_leg(pinned):
        if pinned:
            hs = [{k: torch.from_numpy(np.ascontiguousarray(s[k])).pin_memory() for k in keys} for s in sets]
            hX = torch.empty((nb, T, n), dtype=torch.float64).pin_memory()
            hU = torch.empty((nb, T, m), dtype=torch.float64).pin_memory()
            hst, hit = torch.empty(nb, dtype=torch.int32).pin_memory(), torch.empty(nb, dtype=torch.int32).pin_memory()
            ptr = lambda t: C.c_void_p(t.data_ptr())
        else:
            hs = [{k: np.ascontiguousarray(s[k]).copy() for k in keys} for s in sets]
            hX, hU = np.empty((nb, T, n)), np.empty((nb, T, m))
            hst, hit = np.empty(nb, dtype=np.int32), np.empty(nb, dtype=np.int32)
            ptr = lambda t: C.c_void_p(t.ctypes.data)

        def step_host(i):
            d = hs[i % nsets]
            rc = L.fmpc_step(hb._h, C.byref(params), nb, ptr(d["x0"]), ptr(d["x0_pre"]), None, None, None, ptr(d["X0"]), ptr(d["U0"]),
                             ptr(d["nu0"]), ptr(hX), ptr(hU), ptr(hst), ptr(hit), None)
            if rc:
                raise pk.FmpcError(rc, pk.strerror(rc))
        for i in range(2):
            step_host(i)
        barrier(cx)
        t0 = time.perf_counter()
        for i in range(K):
            step_host(i)
        torch.cuda.synchronize(dev)
        return sync_max(cx, time.perf_counter() - t0)

    if hostlegs:
        res["e2e_s"] = host_leg(True)
        res["e2e_pageable_s"] = host_leg(False)
        res["h2d"] = sum(int(np.asarray(sets[0][k]).size) * 8 for k in keys)
        res["d2h"] = nb * T * (n + m) * 8 + nb * 8
    hb.close()
    return res


def roofline_of(cx, res, p, mode):
    F = f_newton(p.n, p.m, p.T)
    flops_per_launch = res["newton_iters"] / res["K"] * F
    achieved = flops_per_launch / (res["kern_ms"] * 1e-3) / 1e12
    return {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA)", "kernel": res["kernel"], "achieved": achieved, "peak": cx.peak,
            "unit": "TFLOP/s", "frac": achieved / cx.peak, "traffic": traffic_of(res["kernel"], res["nb"], mode),
            "traffic_unit": "bytes per launch (ncu dram read + write)",
            "peak_source": "measured live: fmpc_fp64_peak DMMA m8n8k4 %.2f / DFMA %.2f TFLOP/s "
                           "(MEASURED_PEAKS.json has no FP64 entry)" % (cx.peak_dmma, cx.peak_dfma),
            "flops_per_newton_iter": F, "newton_iters_per_launch": res["newton_iters"] / res["K"], "kernel_ms": res["kern_ms"]}


def measure_zmf(cx):
    """zernmodfit (BASELINE.json configs[2]): 2000 frames 128 x 128 onto N = 6, and a 32768-frame bandwidth run."""
    import torch
    import mpc_sensorlessao_b200 as pk
    L, dev = cx.L, cx.dev
    hbm = 6550.4
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {}
    vp = lambda t: C.c_void_p(t.data_ptr())
    for nf in (2000, 32768):
        zf = pk.ZernikeFitter(128, 6, max_frames=8, device=cx.local_rank)
        nm = zf.nmodes
        frames = torch.randn((nf, 128 * 128), dtype=torch.float64, device=dev)
        coef = torch.empty((nf, nm), dtype=torch.float64, device=dev)
        st = torch.cuda.Stream(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(8):
            flush.fill_(it)                               # 256 MB > L2: the frames must come from HBM
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = L.zmf_fit_d(zf._h, nf, vp(frames), vp(coef), C.c_void_p(st.cuda_stream))
            e1.record(st)
            torch.cuda.synchronize(dev)
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        byts = nf * (8 * 128 * 128 + 8 * nm)
        out[f"zernmodfit_{nf}_frames"] = {"kernel_ms": ms, "frames_per_s": nf / ms * 1e3, "l2_policy": "256 MB flush between launches",
                                           "roofline": {"bound": "hbm", "achieved": byts / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                                        "frac": byts / ms / 1e6 / hbm, "bytes_per_frame": byts // nf,
                                                        "fp64_tflops": 2 * nm * zf.npix_in * nf / ms / 1e9}}
        zf.close()
    return out


def measure_c1_batch(cx, nb=4096):
    """BASELINE.json configs[0] shape, batched: the reference's VAR_1 exactly as written (ramp-rate rows + its literal C),
    n = 27, m = 144, T = 10, on the general-structure kernel; one batched solve of nb instances, warm-started."""
    import torch
    import mpc_sensorlessao_b200 as pk
    from mpc_sensorlessao_b200 import synth
    L, dev = cx.L, cx.dev
    p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
    wi = synth.warm_inputs(p, nb, seed=7)
    rs = np.random.RandomState(8)
    u_prev = wi["U0"][:, 0] + 0.01 * rs.randn(nb, p.m)
    hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, du_min=p.du_min, du_max=p.du_max,
                         ramp_rows=True, var1_literal_bug=True, max_batch=nb, device=cx.local_rank)
    params = hb.params(KAPPA, NITERS, 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(x0=wi["x0"], X0=wi["X0"], U0=wi["U0"], nu0=wi["nu0"],
                                                                                u_prev=u_prev).items()}
    X = torch.empty((nb, p.T, p.n), dtype=torch.float64, device=dev)
    U = torch.empty((nb, p.T, p.m), dtype=torch.float64, device=dev)
    st = torch.empty(nb, dtype=torch.int32, device=dev)
    it = torch.empty(nb, dtype=torch.int32, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    stream = torch.cuda.Stream(dev)
    ms = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = L.fmpc_step_d(hb._h, C.byref(params), nb, vp(d["x0"]), None, vp(d["u_prev"]), None, None, vp(d["X0"]), vp(d["U0"]),
                           vp(d["nu0"]), vp(X), vp(U), vp(st), vp(it), C.c_void_p(stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))
        ms.append(e0.elapsed_time(e1))
    kind = hb.kernel_kind
    hb.close()
    t = float(np.median(ms[1:]))
    return {"workload": f"VAR(1) fastMPC as the reference writes it (ramp rows + literal C), n={p.n}, m={p.m}, T={p.T}, {nb} instances, "
                        f"niters={NITERS}, one batched solve (BASELINE.json configs[0] shape, batched)", "kernel_kind": kind,
            "value": nb / t * 1e3, "unit": "solves/s", "kernel_ms": t, "newton_iters_per_solve": float(it.sum().item()) / nb,
            "status_hist": np.bincount(st.cpu().numpy(), minlength=5).tolist()}


def measure_closed_loop(cx, p, nb, K):
    """fmpc_closed_loop: K steps of nb loops in ONE call (aberration sequence up, logs down), MATLAB stream on the device."""
    import mpc_sensorlessao_b200 as pk
    from mpc_sensorlessao_b200 import synth
    a = synth.aberrations(p, nb, K, seed=3)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb, device=cx.local_rank)
    hb.closed_loop(a[:, :3], nu0=None, kappa=KAPPA, niters=NITERS)
    t0 = time.perf_counter()
    out = hb.closed_loop(a, nu0=None, kappa=KAPPA, niters=NITERS)
    wall = time.perf_counter() - t0
    hb.close()
    its = float(out["iters"].sum())
    F = f_newton(p.n, p.m, p.T)
    return {"workload": f"fmpc_closed_loop, {nb} loops x {K} steps in one call (README loop x0 = a + B u_prev, cold start at step 0), "
                        "nu0 = NULL (MATLAB stream generated on the device)",
            "solves_per_s_device": nb * K / out["telapsed"], "solves_per_s_wall": nb * K / wall,
            "newton_iters_per_solve": its / (nb * K), "roofline_frac": its * F / out["telapsed"] / 1e12 / cx.peak}


def run_single_process(args):
    """--single-process: ONE process drives all --gpus devices through the library's own multi-GPU entry points
    (fmpc_multi_step_r: one handle + host thread per device, contiguous shards) -- the path a MATLAB caller gets.
    Host buffers only, so the line's `value` is the end-to-end rate; per-device statistics come back over NCCL."""
    import torch
    import mpc_sensorlessao_b200 as pk
    from mpc_sensorlessao_b200 import synth
    sys.stdout.flush()
    real_stdout = os.dup(1)          # NCCL may print its version banner to stdout: keep stdout to the one JSON line
    os.dup2(2, 1)
    name = args.config if args.config in CONFIGS else "c2"
    N, T, nbc, ub, scaling, _ = CONFIGS[name]
    G = args.gpus
    nb_gpu = args.instances or nbc or 65536 // G
    nb = nb_gpu * G
    p = synth.make_problem(N, T, u_bound=ub)
    n, m = p.n, p.m
    K, W = args.steps, max(args.warmup, 3)
    hm = pk.FastMPCMulti(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T, p.x_min, p.x_max, max_batch=nb, ngpus=G)
    params = hm.params(KAPPA, NITERS, 0)
    KT = PRE + W + K
    noise = np.concatenate([loop_noise(p, nb_gpu, KT + 1, seed=100 + 10 * g) for g in range(G)], axis=0)
    x0s = [torch.empty((nb, n), dtype=torch.float64).pin_memory() for _ in range(KT)]
    hu = torch.empty((nb, m), dtype=torch.float64).pin_memory()
    hs, hi = torch.empty(nb, dtype=torch.int32).pin_memory(), torch.empty(nb, dtype=torch.int32).pin_memory()
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def step(k):
        rc = hm._L.fmpc_multi_step_r(hm._h, C.byref(params), nb, 1 if k == 0 else 0, ptr(x0s[k]), None, None, None, None, None,
                                     ptr(hu), None, None, ptr(hs), ptr(hi), None)
        if rc:
            raise pk.FmpcError(rc, pk.strerror(rc))
    x = noise[:, 0].copy(); xp = np.zeros((nb, n))
    for k in range(KT):                     # pass 0: the plant on the host records the x0 sequence
        x0s[k].numpy()[:] = x
        step(k)
        xn = x @ p.A1.T + xp @ p.A2.T + hu.numpy() @ p.B.T + noise[:, k + 1]
        xp, x = x, xn
    for k in range(PRE + W):
        step(k)
    its = 0
    t0 = time.perf_counter()
    for i in range(K):
        step(PRE + W + i)
        its += int(hi.numpy().sum())
    dt = time.perf_counter() - t0
    st = hm.stats(use_nccl=True)
    launches = hm.launch_count
    hm.close()
    F = f_newton(n, m, T)
    line = {"metric": "fastmpc_solves_per_sec", "value": nb * K / dt, "unit": "solves/s", "n_gpus": G, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_desc(name, p, nb_gpu, G, "loop"),
            "mode": "single process, fmpc_multi_step_r (one handle + host thread per device), pinned host buffers, wall clock",
            "e2e": {"value": nb * K / dt, "unit": "solves/s", "h2d_bytes_per_step": nb * n * 8, "d2h_bytes_per_step": nb * (m * 8 + 8),
                    "api": "fmpc_multi_step_r"},
            "newton_iters_per_solve": its / (nb * K), "aggregate_tflops": its * F / dt / 1e12, "gpu_launches": launches,
            "per_device_stats_last_step": st}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)      # ~1 s of timed steps at C2
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS) + ["zmf", "closed_loop"])
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the side workloads reported under `extra`")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus devices from ONE process through fmpc_multi_* (not under torchrun)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if args.single_process:
        run_single_process(args)
        return

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

    cx = Ctx()
    cx.rank, cx.local_rank, cx.world = rank, local_rank, world
    torch.cuda.set_device(local_rank)
    cx.dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch decides where the pages live
    if world > 1:
        # NCCL prints its version banner to STDOUT when NCCL_DEBUG is VERSION/WARN: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=cx.dev)
    cx.L = load_library()
    # FP64 pipe peaks, measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)
    cx.peak_dmma = pk.fp64_peak(local_rank, 1, 4000)
    cx.peak_dfma = pk.fp64_peak(local_rank, 0, 4000)
    cx.peak = max(cx.peak_dmma, cx.peak_dfma)
    K, W = args.steps, args.warmup

    def emit(line):
        if rank == 0:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)

    if args.config == "zmf":
        z = measure_zmf(cx)
        r = z["zernmodfit_2000_frames"]
        emit({"metric": "zernmodfit_frames_per_sec", "value": r["frames_per_s"], "unit": "frames/s", "n_gpus": 1, "steps": 6, "warmup": 2,
              "ms_per_step": r["kernel_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic", "config": {"workload": "zernmodfit of 2000 synthetic 128x128 frames onto N=6 (BASELINE.json configs[2])"},
              "roofline": r["roofline"], "extra": z, "gpu_launches": 16})
        return
    name = args.config if args.config in CONFIGS else "c2"
    N, T, nbc, ub, scaling, _ = CONFIGS[name]
    nb = args.instances or nbc or 65536 // world
    p = synth.make_problem(N, T, u_bound=ub)
    if args.config == "closed_loop":
        cl = measure_closed_loop(cx, p, nb, max(K, 8))
        emit({"metric": "fastmpc_solves_per_sec", "value": cl["solves_per_s_device"], "unit": "solves/s", "n_gpus": 1, "steps": max(K, 8),
              "warmup": 3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": cl["workload"]}, "extra": {"closed_loop": cl}})
        return
    mode = "sets" if name == "c2_tight" else "loop"
    res = (measure_loop if mode == "loop" else measure_sets)(cx, name, p, nb, K, W)
    full = measure_sets(cx, name, p, nb, max(4, min(K, 20)), 3) if (mode == "loop" and name == "c2") else None

    # ---- gather statistics (NCCL carries only this) ----
    stats = torch.tensor([float(res["newton_iters"]), float(res["launches"])], dtype=torch.float64, device=cx.dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    newton_iters_all, launches_all = float(stats[0].item()), int(stats[1].item())

    line = None
    if rank == 0:
        F = f_newton(p.n, p.m, p.T)
        solves = nb * K * world
        line = {
            "metric": "fastmpc_solves_per_sec", "value": solves / (res["total_ms"] * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": res["total_ms"] / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_desc(name, p, nb, world, mode),
            "clocks": res["clocks"],
            "e2e": {"value": solves / res["e2e_s"], "unit": "solves/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "api": ("fmpc_step_r (C-ABI, pinned host buffers): x0 in, U(:,0) + status + iters out; dual start = the handle's MATLAB "
                            "stream, generated on the device" if mode == "loop" else "fmpc_step (C-ABI, pinned host buffers)"),
                    "numa_bound_cpus": numa},
            "e2e_pageable": {"value": solves / res["e2e_pageable_s"], "unit": "solves/s",
                             "api": "same call with ordinary malloc'd (pageable) host buffers -- what mxGetPr hands the MEX gateway"},
            "gpu_launches": launches_all,
            "roofline": roofline_of(cx, res, p, mode),
            "newton_iters_per_solve": newton_iters_all / solves, "status_hist": res["status_hist"], "iters_hist": res["iters_hist"],
            "aggregate_tflops": newton_iters_all * F / (res["total_ms"] * 1e-3) / 1e12,
        }
        if "rms_x0" in res:
            line["loop_state"] = {"rms_x0_rad": res["rms_x0"], "max_abs_u0": res["u0_max"]}
            line["config"]["workspace_mb"] = res["ws_mb"]
        if full is not None:
            sf = nb * full["K"] * world
            line["e2e_full"] = {"value": sf / full["e2e_s"], "unit": "solves/s", "h2d_bytes_per_step": full["h2d"],
                                "d2h_bytes_per_step": full["d2h"],
                                "api": "fmpc_step (C-ABI, pinned host buffers): whole warm start (X0, U0) and nu0 in, whole horizon (X, U) out"}
            line["e2e_full_pageable"] = {"value": sf / full["e2e_pageable_s"], "unit": "solves/s",
                                         "api": "fmpc_step with pageable host buffers (library-side pinned staging ring + copy threads)"}
            line["value_full_surface_inputs"] = {"value": sf / (full["total_ms"] * 1e-3), "unit": "solves/s",
                                                 "roofline_frac": roofline_of(cx, full, p, "sets")["frac"],
                                                 "newton_iters_per_solve": full["newton_iters"] / (nb * full["K"]),
                                                 "what": "round-1 headline workload: fmpc_step_d on rotating warm input sets in HBM"}
    if world == 1 and rank == 0:
        extra = {}
        if not args.no_extra and name == "c2":
            try:
                pt = synth.make_problem(6, 20, u_bound=1.0)
                rt = measure_sets(cx, "c2_tight", pt, 4096, 4, 3, hostlegs=False)
                extra["c2_tight"] = {"workload": workload_desc("c2_tight", pt, 4096, 1, "sets")["workload"],
                                     "value": 4096 * rt["K"] / (rt["total_ms"] * 1e-3), "unit": "solves/s",
                                     "newton_iters_per_solve": rt["newton_iters"] / (4096 * rt["K"]), "iters_hist": rt["iters_hist"],
                                     "status_hist": rt["status_hist"], "roofline": roofline_of(cx, rt, pt, "sets")}
                from oracle import fmpc_ref
                wi = rt["sets"][0]
                z0 = np.concatenate([wi["U0"][:64], wi["X0"][:64]], axis=2).reshape(64, -1)
                o = fmpc_ref.solve_batch(pt.A1, pt.A2, pt.B, pt.Q, pt.R, pt.Qf, pt.u_min, pt.u_max, KAPPA, NITERS, wi["x0"][:64].T,
                                         wi["x0_pre"][:64].T, None, z0.T, wi["nu0"][:64].T)
                extra["c2_tight"]["halvings_hist_oracle_64_instances"] = np.bincount(np.asarray(o["halvings"]).reshape(-1)).tolist()
                del rt
                p5 = synth.make_problem(10, 30)
                r5 = measure_sets(cx, "c5", p5, 16384, 3, 3, nsets=1, hostlegs=False)
                extra["c5"] = {"workload": workload_desc("c5", p5, 16384, 1, "sets")["workload"],
                               "value": 16384 * r5["K"] / (r5["total_ms"] * 1e-3), "unit": "solves/s",
                               "newton_iters_per_solve": r5["newton_iters"] / (16384 * r5["K"]), "roofline": roofline_of(cx, r5, p5, "sets")}
                del r5
                extra["c1_batch"] = measure_c1_batch(cx)
                extra.update(measure_zmf(cx))
                extra["closed_loop"] = measure_closed_loop(cx, p, 4096, 20)
            except Exception as e:          # a side workload must never cost the headline line
                extra["error"] = repr(e)
        if extra:
            line["extra"] = extra
        if not args.no_cpu_baseline:
            if ORIG_AFFINITY:
                os.sched_setaffinity(0, ORIG_AFFINITY)      # the CPU baseline uses every host thread again
            cores = host_threads()
            if mode == "loop":
                nsample = min(nb, 2048 if p.n <= 32 else 256)
                r0, _, _, dt0 = cpu_loop_rate(p, nsample, 1, 0, cores)
                reps = max(3, min(60, int(round(12.0 / max(dt0, 1e-3)))))          # ~12 s of CPU work
                rate, it, cnt, dt = cpu_loop_rate(p, nsample, reps, 1, cores)
                sample = (f"{nsample} of the workload's closed loops, {reps} steps ({reps * dt:.1f} s of solves), structured C/OpenMP port "
                          "(oracle/fmpc_ref.c) on every host thread; the MATLAB reference cannot run on this box")
            else:
                r0, _, _, dt0 = cpu_port_rate(p, res["sets"], nb, nthreads=cores, reps=1)
                reps = max(3, min(60, int(round(12.0 / max(dt0, 1e-3)))))
                rate, it, cnt, dt = cpu_port_rate(p, res["sets"], nb, nthreads=cores, reps=reps)
                sample = f"all {cnt} instances of one step, {reps} repetitions ({reps * dt:.1f} s), structured C/OpenMP port"
            line["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
