#!/usr/bin/env python
"""Benchmark: env-steps/sec of Shadow-hand cube reorientation (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this build (CUDA kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU arm on the box's host cores

A "step" is one `Environment.step` over the whole batch (16384 envs per GPU, each 5 mj_steps
+ task pre/post kernels, auto-reset included).  `value` is whole-job env-steps/sec with actions
resident in HBM; `e2e` is the same through the public API with HOST actions (pinned H2D every
step) and the TimeStep read back to the host.  Prints ONE JSON line on rank 0.

The CPU arm is the in-repo FP64 restatement of mj_step (oracle/, "port"): the reference's own
engine (un-vendored `mujoco` + `dm_control`) is not installable here — see DESIGN.md.  It is
never presented as a MuJoCo number.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_ENVS_PER_GPU = 16384          # BASELINE.json configs[2] / configs[4] (131072 over 8 GPUs)
B_ENV = 1364                      # algorithmic bytes per env-step, SURVEY.md §8(d) / BASELINE.md §3
B_MJ = 808                        # algorithmic bytes per mj_step (state round trip)
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 16384 envs from the `ncu --set full` capture summarised in
# profiles/r02z_summary.md (bench.py cannot run under ncu; the figure is only reported at that batch size).  k_dynamics = the dense
# solver kernel plus the structured kernel's contact-free form, which run side by side on two streams.
NCU_TRAFFIC_BYTES = {"k_dynamics": 55.54e6 + 2.05e6 + 30.08e6, "k_tree": 46.55e6 + 68.79e6, "k_collide": 35.76e6, "k_narrow": 12.28e6}
METRIC = "env-steps/sec, Shadow-hand cube reorient"
UNIT = "env-steps/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region: NVML polled every 5 ms from a thread
    (the region is ~100 ms at the default step count), `nvidia-smi -lms` as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.lines, self.gpu = None, [], gpu_index
        self.nvml, self.handle, self.thread, self.stop_flag = None, None, None, False
        self.sm, self.smax, self.reasons = [], [], set()

    def _nvml_loop(self):
        N = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(N, a, None) or getattr(N, b, 0)) for n, a, b in names]
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.handle, N.NVML_CLOCK_SM)))
                r = int(get_reasons(self.handle))
                for n, b in bits:
                    if b and (r & int(b)):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as N
            import torch
            N.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                self.handle = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                self.handle = N.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = [float(N.nvmlDeviceGetMaxClockInfo(self.handle, N.NVML_CLOCK_SM))]
            self.nvml = N
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.smax) if self.smax else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def cpu_arm(steps: int, warmup: int, bounded_seconds: float = 14.0):
    """Times the CPU restatement of the path on all host cores (bounded sample)."""
    import numpy as np
    from dexterity_b200.model import scenes
    from oracle import oracle as orc
    from tests import statelib
    compiled = scenes.load_compiled("reorient_shadow")
    om = orc.OracleModel(compiled)
    cores = os.cpu_count() or 1
    n = 32 * cores
    rng = np.random.RandomState(12345)
    qpos = np.stack([statelib.reorient_reset(compiled, rng) for _ in range(n)])
    qvel = np.zeros((n, compiled.nv)); warm = np.zeros((n, compiled.nv))
    lo, hi = compiled.ACT_CTRLRANGE[:, 0], compiled.ACT_CTRLRANGE[:, 1]
    per_step_s = []
    total = warmup + steps
    t_begin = time.perf_counter()
    done = 0
    for k in range(total):
        ctrl = rng.uniform(lo, hi, size=(n, compiled.nu))
        t0 = time.perf_counter()
        om.step_batch(qpos, qvel, ctrl, warm, nsub=5, nthreads=cores)
        dt = time.perf_counter() - t0
        if k >= warmup:
            per_step_s.append(dt)
        done = k + 1
        if time.perf_counter() - t_begin > bounded_seconds and len(per_step_s) >= 1:
            break
    secs = float(np.sum(per_step_s))
    value = n * len(per_step_s) / secs
    sample = (f"{n} envs x {len(per_step_s)} control steps (5 mj_steps + refresh each) from the reorient reset "
              f"distribution, uniform random ctrl, {cores} host threads, FP64 restatement of mj_step (not MuJoCo)")
    return value, cores, sample, 1e3 * secs / max(1, len(per_step_s)), len(per_step_s), done


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    value, cores, sample, ms, nsteps, _ = cpu_arm(args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": nsteps,
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "reorient (Shadow hand + cube, full contact solve), CPU arm: bounded sample",
                   "envs": 32 * cores, "substeps": 5},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_gpu(args):
    import numpy as np
    import torch
    from dexterity_b200 import backend, envs
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on STDOUT; rank 0 must print ONE JSON line
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.num_envs
    env = envs.load("reorient", "state_dense", seed=12345 + rank, num_envs=n, device=local)
    spec = env.action_spec()
    lo = torch.as_tensor(spec.minimum, device=dev); hi = torch.as_tensor(spec.maximum, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    nu = spec.shape[0]
    pool = [lo + (hi - lo) * torch.rand(n, nu, device=dev, generator=gen) for _ in range(8)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    lib = backend.lib()
    import ctypes
    lib.dex_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.dex_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    env.reset()
    for k in range(args.warmup):
        env.step(pool[k % len(pool)])
    barrier()

    # ---------------- device-resident throughput (`value`) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = env.physics.launch_count
    lib.dex_set_profiling(env.physics.ptr, 1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))                      # L2 flush between timed iterations (untimed)
        env.action_buffer.copy_(pool[k % len(pool)])
        ev[k][0].record()
        env.step_bound()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = env.physics.launch_count - launches0
    ms4 = (ctypes.c_double * 10)(); cnt4 = (ctypes.c_int64 * 10)()
    backend._check(lib.dex_get_kernel_times(env.physics.ptr, ms4, cnt4, 1))
    lib.dex_set_profiling(env.physics.ptr, 0)
    clocks = sampler.stop() if rank == 0 else None

    # the same steps back to back without the L2 flush (the state and scratch one step touches, ~160 MB, exceed the 126 MB L2
    # anyway); reported beside `value`, which stays the flushed, more conservative figure
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for k in range(args.steps):
        env.action_buffer.copy_(pool[k % len(pool)])
        env.step_bound()
    u1.record()
    barrier()
    unflushed_ms = torch.tensor([u0.elapsed_time(u1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(unflushed_ms, op=torch.distributed.ReduceOp.MAX)
    unflushed_ms = float(unflushed_ms.item())

    # ---------------- end to end through the public API with host buffers ----------------
    host_actions = [p.cpu().pin_memory() for p in pool[:4]]
    # the observation rows as they lie in device memory (128 floats, 123 in use): one contiguous DMA instead of a strided copy
    obs_dev = env.observation_buffer if os.environ.get("DEX_BENCH_STRIDED_OBS") else env._obs
    obs_host = torch.empty(n, obs_dev.shape[1], dtype=torch.float32).pin_memory()
    rew_host = torch.empty(n, dtype=torch.float32).pin_memory()
    dis_host = torch.empty(n, dtype=torch.float32).pin_memory()
    typ_host = torch.empty(n, dtype=torch.int32).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        a = host_actions[k % len(host_actions)].to(dev, non_blocking=True)      # H2D of this step's inputs
        ts = env.step(a)
        obs_host.copy_(obs_dev, non_blocking=True)                              # D2H of the TimeStep
        rew_host.copy_(ts.reward, non_blocking=True); dis_host.copy_(ts.discount, non_blocking=True)
        typ_host.copy_(ts.step_type, non_blocking=True)
        torch.cuda.current_stream().synchronize()                               # the caller needs the result
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = t.tolist()
    stats = env.reduce_episode_stats()          # the path's only collective (8 doubles, NCCL)
    status = env.physics.status
    nbad = int((status & 1).ne(0).sum().item())

    if rank == 0:
        total_envs = n * world
        value = total_envs * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_envs * args.steps / (e2e_ms_max * 1e-3)
        peak, peak_src = measured_peaks()
        # k_dynamics = dense warp-per-env solver (envs with contacts), k_dynamics_arrow = structured eight-lanes-per-env solver
        # (contact-free envs); the two run concurrently on two streams, so their event times overlap
        names = ["task_pre", "k_collide", "k_dynamics", "task_post", "k_tree", "k_narrow", "k_compact", "k_dynamics_arrow_full",
                 "k_place", "k_dynamics_arrow"]
        per = {names[c]: (ms4[c] / max(1, cnt4[c]), cnt4[c]) for c in range(10) if cnt4[c] or c < 7}
        dom = max(("k_tree", "k_collide", "k_narrow", "k_dynamics"), key=lambda k_: per[k_][0])
        dom_ms = per[dom][0]
        achieved = n * B_MJ / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        cpu_value, cores, sample, _, _, _ = cpu_arm(steps=1000, warmup=1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "reorient (Shadow hand + cube, full contact solve), 16384 envs per GPU "
                                   "(BASELINE.json configs[2]; configs[4] at N=8)",
                       "envs_per_gpu": n, "substeps": 5, "actions": "uniform random in ctrlrange",
                       "auto_reset": True, "l2": "flushed between timed steps (256 MiB write, untimed)"},
            "mj_steps_per_sec": 5 * value,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * nu * 4,
                    "d2h_bytes_per_step": n * (obs_dev.shape[1] + 3) * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(dom) if n == NUM_ENVS_PER_GPU else None, "kernel": dom,
                         "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": n * B_MJ, "peak_source": peak_src,
                         "env_step_frac": (value / world) * B_ENV / 1e9 / peak,
                         "note": "compute/latency-bound path: HBM fraction is reported for comparability"},
            "kernel_ms": {k_: round(v[0], 4) for k_, v in per.items()},
            "no_l2_flush": {"value": n * world / unflushed_ms * 1e3, "unit": UNIT, "ms_per_step": unflushed_ms,
                            "note": "same steps back to back without the flush (includes the action copy on the device)"},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "clocks": clocks,
            "episode_stats": stats, "physics_errors_in_batch": nbad, "wall_s": wall,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


_JSON_FD = None


def capture_stdout():
    """The driver reads ONE JSON line from stdout, but libraries write there too (NCCL prints its version banner on
    stdout at the first collective): point fd 1 at stderr for the whole run and keep the real stdout for emit()."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-envs", type=int, default=NUM_ENVS_PER_GPU)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
