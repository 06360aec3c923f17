#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on B200: OSC control evaluations / second, batched UR5 6-DOF.

One "step" = one pass of the hot path (the fused OSC kernel: chain walk -> J, M, g, C dq -> solves -> u) over
one batch of B synthetic joint states per GPU (B = 65 536, fp64: the UR5 configuration of BASELINE.json,
configs[1], driven through OSC.generate with use_C so that {J, M, g, c_forces} are all on the path).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line on rank 0)
  python bench.py --impl reference [...]                         # the reference's CPU path on the host cores

Under torchrun (N > 1) every rank owns its own B states (weak scaling, no data-path collective; --allgather adds
the optional NCCL all-gather of the control outputs).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  L2: the steps rotate over a ring of input/output buffer sets larger
than the 126 MB L2, so every step's inputs come from HBM.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "OSC control evals/sec (batched UR5 6-DOF)"
UNIT = "evals/s"
B_PER_GPU = 65536
OSC_KW = dict(kp=10.0, ctrlr_dof=[True] * 6, use_C=True)  # 6 controlled DOF, gravity + Coriolis compensation
WORKLOAD = "ur5_osc_6dof_useC_fp64_B65536_per_gpu"


def synth(B, n, seed, dtype=np.float64):
    """seeded synthetic states as examples/timing_plots.py:18-20: q~U(0,2pi), dq~U(0,5), target~U(-1,1)"""
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, 2 * np.pi, (B, n)).astype(dtype), rng.uniform(0, 5, (B, n)).astype(dtype),
            rng.uniform(-1, 1, (B, 6)).astype(dtype))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.

    Primary source: NVML in this process (the library nvidia-smi itself reads), polled every 2 ms by a thread — it has
    no start-up latency, so even a 50 ms timed region gets samples.  `nvidia-smi --query-gpu=... -lms 10` runs beside
    it as a second source (it needs ~100 ms to print its first line); the two are merged."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, uuid=None):
        import threading

        self.sm, self.mx, self.reasons, self.src = [], [], set(), []
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self._stop = threading.Event()
        self._thread = None
        period = os.environ.get("ABRB_BENCH_CLOCK_MS", "10")  # tuning knob: "off" disables the sampler (A/B only)
        if period == "off":
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (f"GPU-{uuid}", str(uuid)):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons",
                                 getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        if reasons_fn is not None:
                            mask = int(reasons_fn(h))
                            for bit, name in self.BITS.items():
                                if mask & bit:
                                    self.reasons.add(name)
                    except Exception:
                        pass
                    self._stop.wait(0.002)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            self.src.append("nvml")
        except Exception:
            self._thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       period, "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": []}
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1)
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
            n_smi = 0
            for line in self.f.read().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    self.sm.append(float(c[1]))
                    self.mx.append(float(c[2]))
                    n_smi += 1
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            if n_smi:
                self.src.append("nvidia-smi")
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)) if self.mx else None,
                       reasons=sorted(self.reasons), samples=len(self.sm), source=self.src)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------------- reference arm
def ref_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libabrref_ur5.so")
    return C.CDLL(p) if os.path.exists(p) else None


class OscCfg(C.Structure):
    _fields_ = [("n", C.c_int), ("kp", C.c_double), ("ko", C.c_double), ("kv", C.c_double), ("use_vmax", C.c_int),
                ("vmax", C.c_double * 2), ("dof", C.c_int * 6), ("use_g", C.c_int), ("use_C", C.c_int),
                ("alg", C.c_int), ("damp_kv", C.c_double), ("use_rest", C.c_int), ("rest_kp", C.c_double),
                ("rest_kv", C.c_double), ("rest", C.c_double * 8), ("rest_mask", C.c_int * 8)]


def osc_cfg():
    c = OscCfg()
    c.n = 6
    c.kp = c.ko = OSC_KW["kp"]
    c.kv = float(np.sqrt(c.kp + c.ko))
    for r in range(6):
        c.dof[r] = 1
    c.use_g, c.use_C, c.alg, c.damp_kv = 1, 1, 0, -1.0
    return c


def cpu_reference_run(n_evals, repeats=3):
    """OSC evals/s of the reference's CPU path on all host cores.

    kind "reference": the reference's own SymPy-generated C for J/Tx/M/g/C/R (oracle/_ref, compiled from where the
    reference wrote it) + oracle/c/osc_cpu.c for the NumPy half, pthreads over states.
    kind "port" (fallback when oracle/_ref is absent): the NumPy oracle, one core."""
    q, dq, target = synth(n_evals, 6, 123)
    lib = ref_lib()
    if lib is not None:
        lib.ref_max_threads.restype = C.c_int
        online = int(lib.ref_max_threads())
        try:
            online = min(online, len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        cfg = osc_cfg()
        u = np.empty((n_evals, 6))

        def run(nthreads, count):
            a = (C.byref(cfg), q.ctypes.data_as(C.c_void_p), dq.ctypes.data_as(C.c_void_p),
                 target.ctypes.data_as(C.c_void_p), C.c_long(count), u.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
            t0 = time.perf_counter()
            lib.ref_ur5_osc_batch(*a)
            return time.perf_counter() - t0

        # the box may expose more hardware threads than it lets us use: pick the best thread count on a short probe
        probe = min(n_evals, 400_000)
        run(online, probe)  # warm (page in)
        cands = sorted({c for c in (1, online // 8, online // 4, online // 2, online) if c >= 1})
        rates = {c: probe / min(run(c, probe) for _ in range(2)) for c in cands}
        cores = max(rates, key=rates.get)
        best = min(run(cores, n_evals) for _ in range(repeats))
        return dict(value=n_evals / best, unit=UNIT, cores=cores, kind="reference",
                    threads_probe={str(k): round(v) for k, v in rates.items()}, hw_threads_online=online,
                    sample=f"{n_evals} UR5 OSC evals (same controller, seeded states), best of {repeats}, "
                           "reference-generated C for J/Tx/M/g/C/R + C restatement of the NumPy half, pthreads"), u, (q, dq, target)
    from oracle import osc_oracle

    n_small = min(n_evals, 300)
    case = dict(arm="ur5", osc=OSC_KW)
    t0 = time.perf_counter()
    u, _ = osc_oracle.run_case(case, q[:n_small], dq[:n_small], target[:n_small])
    dt = time.perf_counter() - t0
    return dict(value=n_small / dt, unit=UNIT, cores=1, kind="port",
                sample=f"{n_small} UR5 OSC evals, NumPy oracle (oracle/_ref absent)"), u, (q[:n_small], dq[:n_small], target[:n_small])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step a bounded sample of the workload: four batches per call so that starting the worker threads (once per
    # call in oracle/c/ref_ur5_driver.c) stays a small part of it
    n_step = 4 * B_PER_GPU
    base, _, _ = cpu_reference_run(n_step, repeats=1)
    lib = ref_lib()
    q, dq, target = synth(n_step, 6, 123)
    times = []
    if lib is not None:
        cfg = osc_cfg()
        u = np.empty((n_step, 6))
        a = (C.byref(cfg), q.ctypes.data_as(C.c_void_p), dq.ctypes.data_as(C.c_void_p),
             target.ctypes.data_as(C.c_void_p), C.c_long(n_step), u.ctypes.data_as(C.c_void_p), C.c_int(base["cores"]))
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            lib.ref_ur5_osc_batch(*a)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        per_step = float(np.median(times))  # the host cores are shared with other tenants of the box
        value = n_step / per_step
    else:
        value, per_step = base["value"], n_step / base["value"]
    base["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n_step, "arm": "ur5", "osc": "kp=10, ctrlr_dof=[T]*6, use_C, use_g"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- our arm
def time_kernel(fn, n_launch, torch, sets):
    """average device time of one launch, CUDA events on the current stream, rotating buffer sets"""
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_launch):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n_launch


def run_ours(args):
    import torch
    import torch.distributed as dist

    from abr_control_b200 import _lib
    from abr_control_b200.arms import jaco2, ur5
    from abr_control_b200.controllers import OSC, Damping

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, n = B_PER_GPU, 6
    rc = ur5.Config()
    ctrlr = OSC(rc, **OSC_KW)
    L = _lib.lib()

    # ring of buffer sets > L2 (126 MB): each set = q, dq, target in (144 B/state) + u out (48 B/state)
    n_sets = 48
    sets = []
    for s in range(n_sets):
        q, dq, tg = synth(B, n, 1000 * rank + s)
        sets.append(tuple(torch.as_tensor(a, device=dev) for a in (q, dq, tg)))
    ring_mb = n_sets * B * (18 + 6) * 8 / 1e6
    gather_buf = torch.empty((world * B, n), dtype=torch.float64, device=dev) if (args.allgather and world > 1) else None

    outs = [torch.empty((B, n), dtype=torch.float64, device=dev) for _ in range(n_sets)]

    def step(i):
        q, dq, tg = sets[i % n_sets]
        u = ctrlr.generate_into(q, dq, tg, outs[i % n_sets])  # public allocation-free API: one ctypes call
        if gather_buf is not None:
            dist.all_gather_into_tensor(gather_buf, u)
        return u

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    fence()
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if rank == 0 else None
    n0 = L.abrb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    fence()
    launches = L.abrb_launch_count() - n0
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    t = float(elapsed.item())
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / t

    # ---- end to end through the public API with pinned HOST buffers (H2D + kernel + D2H inside every call)
    hq, hdq, htg = (torch.as_tensor(a).pin_memory() for a in synth(B, n, 77 + rank))
    nq, ndq, ntg = hq.numpy(), hdq.numpy(), htg.numpy()
    ctrlr.record_training_signal = False  # the side channel for DynamicsAdaptation is not part of the metric
    for _ in range(3):
        ctrlr.generate(nq, ndq, ntg)
    fence()
    # nine short blocks, median block: host-side copies share the box with whatever else runs on its cores
    e2e_blocks = 9
    per_block = max(2, min(args.steps, 200) // e2e_blocks)
    e2e_steps = e2e_blocks * per_block
    block_t = []
    for _ in range(e2e_blocks):
        t0 = time.perf_counter()
        for _ in range(per_block):
            u_host = ctrlr.generate(nq, ndq, ntg)
        block_t.append(time.perf_counter() - t0)
    t_e2e = torch.tensor(sorted(block_t)[e2e_blocks // 2:e2e_blocks // 2 + 1], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * per_block / float(t_e2e.item())
    e2e_spread = [B * per_block / t for t in (max(block_t), min(block_t))]  # this rank's slowest / fastest block

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    bytes_per_state = (6 + 6 + 6) * 8 + 6 * 8  # q, dq, target in; u out (fp64)
    kernel_s = t / args.steps if gather_buf is None else None
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            for k, v in json.load(fh).items():
                if k.startswith("osc:osc_kernel<double, 6"):
                    traffic = v["dram_mb_per_launch"] * 1e6  # bytes per launch (profiles/<tag>_osc.txt)

    extra = {}
    quick = bool(os.environ.get("ABRB_BENCH_QUICK"))  # tuning knob: headline + e2e only
    if world == 1 and not quick:
        # the other single-GPU kernels, same timing discipline (explain the headline; not bench lines themselves)
        def rbd_sets(want, dtype):
            out = []
            for s in range(24):
                q, dq, _ = synth(B, n, 5000 + s, np.float64 if dtype == torch.float64 else np.float32)
                out.append((torch.as_tensor(q, device=dev), torch.as_tensor(dq, device=dev)))
            return out

        shp = dict(J=(B, 6, n), M=(B, n, n), g=(B, n), C=(B, n, n))

        def mk_rbd(rcx, want, dtype=torch.float64):
            o = {k: torch.empty(shp[k], dtype=dtype, device=dev) for k in want}
            return lambda s: rcx.eval_into(s[0], s[1], o)

        rs = rbd_sets(None, torch.float64)
        for key, want, nbytes in (("rbd_ur5_JMgC_f64", ("J", "M", "g", "C"), 1008), ("rbd_ur5_JMg_f64", ("J", "M", "g"), 672)):
            dt = time_kernel(mk_rbd(rc, want), 200, torch, rs)
            extra[key] = {"states_per_s": B / dt, "us_per_launch": dt * 1e6, "bytes_per_state": nbytes,
                          "achieved_gbs": B * nbytes / dt / 1e9, "frac_hbm": B * nbytes / dt / 1e9 / hbm_peak, "B": B}
        rs32 = rbd_sets(None, torch.float32)
        dt = time_kernel(mk_rbd(rc, ("J", "M", "g"), torch.float32), 200, torch, rs32)
        extra["rbd_ur5_JMg_f32"] = {"states_per_s": B / dt, "us_per_launch": dt * 1e6, "bytes_per_state": 336,
                                    "achieved_gbs": B * 336 / dt / 1e9, "frac_hbm": B * 336 / dt / 1e9 / hbm_peak, "B": B}
        # BASELINE config 3: Jaco2 OSC 5-DOF + Damping, fp32, B = 262144
        B3 = 262144
        rc3 = jaco2.Config()
        c3 = OSC(rc3, kp=200, ctrlr_dof=[True] * 5 + [False], null_controllers=[Damping(rc3, kv=10)])
        s3 = []
        for s in range(16):
            q, dq, tg = synth(B3, 6, 9000 + s, np.float32)
            s3.append(tuple(torch.as_tensor(a, device=dev) for a in (q, dq, tg)))
        u3 = torch.empty((B3, 6), dtype=torch.float32, device=dev)
        dt = time_kernel(lambda s: c3.generate_into(s[0], s[1], s[2], u3), 100, torch, s3)
        extra["osc_jaco2_cfg3_f32_B262144"] = {"evals_per_s": B3 / dt, "us_per_launch": dt * 1e6, "bytes_per_state": 96,
                                               "achieved_gbs": B3 * 96 / dt / 1e9, "frac_hbm": B3 * 96 / dt / 1e9 / hbm_peak}
        # BASELINE config 5 (per-GPU share): Jaco2 OSC xyz + vmax + AvoidObstacles(1 obstacle) + Damping, fp32, B = 131072
        from abr_control_b200.controllers import AvoidObstacles
        B5 = 131072
        c5 = OSC(rc3, kp=200, vmax=[0.5, 0], ctrlr_dof=[True, True, True, False, False, False],
                 null_controllers=[AvoidObstacles(rc3, obstacles=[[0.09596, -0.2661, 0.64204, 0.05]], threshold=0.2),
                                   Damping(rc3, kv=10)])
        s5 = [tuple(t_[:B5].contiguous() for t_ in s) for s in s3[:8]]
        u5 = torch.empty((B5, 6), dtype=torch.float32, device=dev)
        dt = time_kernel(lambda s: c5.generate_into(s[0], s[1], s[2], u5), 50, torch, s5)
        extra["osc_jaco2_cfg5_avoid_f32_B131072"] = {"evals_per_s": B5 / dt, "us_per_launch": dt * 1e6, "bytes_per_state": 96}
        # BASELINE config 4: UR5 OSC(kp=10) closed-loop rollout, 4096 trajectories x 128 steps, dt = 1e-3 (one launch)
        c4 = OSC(rc, kp=10.0)
        q4, dq4, tg4 = (torch.as_tensor(a, device=dev) for a in synth(4096, 6, 4242))
        dq4 = dq4 * 0.1
        for _ in range(2):
            c4.rollout(q4, dq4, tg4, steps=128, dt=1e-3, record=())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            c4.rollout(q4, dq4, tg4, steps=128, dt=1e-3, record=())
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3 / 5
        extra["rollout_ur5_cfg4_f64_4096x128"] = {"osc_evals_per_s": 4096 * 128 / dt, "ms_per_rollout": dt * 1e3,
                                                   "us_per_step": dt / 128 * 1e6, "note": "latency bound: 32 warps per 148 SMs"}
        ctrl32 = OSC(ur5.Config(), **OSC_KW)
        s32 = [tuple(t_.float() for t_ in s) for s in sets[:24]]
        u32b = torch.empty((B, 6), dtype=torch.float32, device=dev)
        dt = time_kernel(lambda s: ctrl32.generate_into(s[0], s[1], s[2], u32b), 200, torch, s32)
        extra["osc_ur5_6dof_f32_B65536"] = {"evals_per_s": B / dt, "us_per_launch": dt * 1e6, "bytes_per_state": 96}

    cpu = None
    if world == 1 and not quick:
        cpu, u_cpu, (cq, cdq, ctg) = cpu_reference_run(4_000_000 if ref_lib() is not None else 300)
        # parity spot check in the same run: first 4096 states of the CPU sample against the GPU path
        m = min(4096, len(cq))
        ug = ctrlr.generate(cq[:m], cdq[:m], ctg[:m])
        rel = np.abs(ug - u_cpu[:m]).max(axis=1) / np.abs(u_cpu[:m]).max(axis=1)
        cpu["parity_vs_gpu_median_rel"] = float(np.median(rel))
        cpu["parity_vs_gpu_p99_rel"] = float(np.quantile(rel, 0.99))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "ur5", "batch_per_gpu": B, "global_batch": world * B,
                   "osc": "kp=10, ctrlr_dof=[T]*6, use_C=True, use_g=True, orientation_algorithm=0",
                   "parallelism": f"batch sharded over {world} GPU(s), no data-path collective" +
                                  (" + NCCL all-gather of u" if gather_buf is not None else ""),
                   "l2": f"inputs rotate over a ring of {n_sets} buffer sets ({ring_mb:.0f} MB > 126 MB L2)"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * 18 * 8), "d2h_bytes_per_step": int(B * 6 * 8),
                "steps": e2e_steps, "blocks": e2e_blocks, "block_range": e2e_spread,
                "how": "OSC.generate(q, dq, target) on pinned host NumPy buffers -> abrb_osc_generate_host_f64 (H2D of q, dq, "
                       "target / kernel / D2H of u into a page-locked result, stream sync inside every call); wall clock, "
                       "median of the nine blocks"},
        "roofline": {"bound": "hbm", "achieved": (B * bytes_per_state / kernel_s / 1e9) if kernel_s else None,
                     "peak": hbm_peak, "unit": "GB/s",
                     "frac": (B * bytes_per_state / kernel_s / 1e9 / hbm_peak) if kernel_s else None, "traffic": traffic,
                     "kernel": "osc_kernel<double,6,ORTHO,KD=6>", "algorithmic_bytes_per_state": bytes_per_state,
                     "peak_source": peak_src,
                     "note": "192 B/state against ~10^4 fp64 flops/state: this kernel is FP64-pipe bound, not HBM bound "
                             "(SURVEY.md S8d); the HBM-bound figures are the rbd_* kernels under 'kernels'"},
        "kernels": extra,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # rank 0 must print exactly ONE line on stdout (the JSON); libraries (NCCL's version banner, ...) write there too,
    # so everything else is sent to stderr and the JSON line goes to the saved descriptor
    global print
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):  # noqa: A001
        k.setdefault("file", real_out)
        k.setdefault("flush", True)
        _print(*a, **k)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--allgather", action="store_true", help="all-gather u across ranks every step (config 5 style)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
