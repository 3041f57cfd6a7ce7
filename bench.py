#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json on N GPUs of one node.

Workload (BASELINE config 3): Dupire local-vol up-and-out barrier call (K 120, B 150,
3y, weekly monitoring, smoothing 1 %), 30 x 36 local-vol surface, 2^20 Sobol paths x 156 steps,
value + full AAD risk (delta + 1080 vegas) = dupireAADRisk.  One "step" = one such pricing-and-risk
pass.  Metric: paths/sec including the full AAD risk (whole job, all GPUs), fp64.

  python bench.py --gpus N --steps K --warmup W            our engine (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N ...             the reference's own CPU path (rank 0 only)

N > 1 is STRONG scaling: the 2^20 paths are split into N disjoint skip-ahead blocks, one per rank
(one process per GPU), and the payoff sums + adjoint vector (1084 doubles) are summed over the ranks inside
the engine's final reduction kernel (peer memory over NVLink: cf_comm_create / cf_comm_connect, CUDA IPC
handles gathered once with torch.distributed), inside the timed region; one NCCL all-reduce per step is the
check and the fallback.
"""
import argparse
import ctypes as C
import json
import os

# rank 0 prints ONE JSON line on stdout: NCCL writes its version banner there at NCCL_DEBUG=VERSION and =WARN
# (init.cc showVersion); drop those two levels before torch loads NCCL (INFO / TRACE, if asked for, are kept)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PATHS = 1 << 20
FLOPS_PER_PATH = 1.63e4          # canonical algorithmic fp64 flops per path incl. AAD
WORKLOAD = "dupire_uoc_barrier_2^20_sobol_paths_x_156_steps_aad_1081_risks"


def config3():
    spots = np.arange(55, 201, 5.0)
    times = np.arange(1, 37) / 12.0
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    return spots, times, vols


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled by a thread through NVML every 5 ms while the timed region
    runs (the timed region is tens of milliseconds: nvidia-smi -lms 100 delivers one sample at best); falls back to
    nvidia-smi when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons = [], [], set()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(self.max_sm)
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Samples taken so far are before the region of interest."""
        self.sm, self.mx, self.reasons = [], [], set()

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(self.NAMES):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def cpu_reference_run(n_paths, repeats=1):
    """The reference's own multi-threaded CPU path (oracle/_ref: main.h dupireAADRisk ->
    mcParallelSimulAAD) on the same inputs; returns (paths/sec best of repeats, threads, value, delta)."""
    from oracle import refapi
    ref = refapi.get()
    threads = ref.start_pool(-1) + 1          # workers + the calling thread, the reference's default
    spots, times, vols = config3()
    ref.put_dupire(100.0, spots, times, vols, 0.25, "bench_dupire")
    ref.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "bench_uoc")
    best, out = 1e30, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = ref.dupire_aad_risk("bench_dupire", "bench_uoc", [1.0, 0.0], 30, 36, n_paths)
        best = min(best, time.perf_counter() - t0)
    return n_paths / best, threads, out[0], out[1], best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = N_PATHS                       # the full job per step: the same configuration as the GPU arm
    for _ in range(max(args.warmup, 1)):
        cpu_reference_run(1 << 14)
    t_tot, pps = 0.0, []
    for _ in range(args.steps):
        p, threads, val, delta, secs = cpu_reference_run(sample)
        pps.append(p); t_tot += secs
    value = sample * args.steps / t_tot
    line = {
        "impl": "reference", "metric": "paths_per_sec_incl_full_aad_risk", "value": value, "unit": "paths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "paths": N_PATHS, "steps_per_path": 156, "surface": "30x36", "rng": "sobol", "risks": 1081,
                   "note": "reference CPU path (mcParallelSimulAAD via dupireAADRisk, oracle/_ref), each step the full 2^20-path job"},
        "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "reference",
                         "sample": f"{sample} paths x 156 steps per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- the other BASELINE configs (parity-test cases at their full sizes): `--config 1|2|4|5` prints the same kind of line
#      for them, through the reference-facing host API; flops per path by SURVEY.md 8d
def _put_config(api, k):
    a = np.arange(10)
    s5 = 100.0 + 5 * a
    spots, times, vols = config3()
    if k in (1, 2):
        (api.put_black_scholes if hasattr(api, "put_black_scholes") else api.put_bs)(100.0, 0.15, False, 0.0, 0.0, "bench_bs")
        if k == 1:
            api.put_european(100.0, 1.0, 1.0, "bench_prd")
        else:
            api.put_barrier(100.0, 120.0, 1.0, 1.0 / 52, 0.01, False, "bench_prd")
        return "bench_bs"
    if k == 4:
        api.put_dupire(100.0, spots, times, vols, 0.25, "bench_dup4")
        api.put_europeans(np.repeat(0.25 * np.arange(1, 13), 60), np.tile(70.5 + np.arange(60), 12), "bench_prd")
        return "bench_dup4"
    api.put_displaced(s5, 0.20 + 0.02 * a, np.where(a % 3 == 0, 0.0, -0.05 * (a % 3)), 0.02, 0.001 * a, [0.5, 1.5],
                      np.full((2, 10), 0.01), np.full((10, 10), 0.5) + 0.5 * np.eye(10), 0.25, "bench_dlm")
    api.put_autocall(s5, 3.0, 12, 1.0, 0.7, 0.10, 0.01, "bench_prd")
    return "bench_dlm"


CONFIGS = {
    # paths, sobol, flops per path, workload, entry point, CPU sample
    1: (1 << 16, True, 1.5e2, "bs_european_2^16_sobol_paths_aad_4_risks", "AADriskOne", 1 << 16),
    2: (1 << 20, True, 5.0e3, "bs_uoc_barrier_2^20_sobol_paths_x_52_steps_aad_4_risks", "AADriskOne", 1 << 18),
    4: (1 << 22, False, 3.8e3, "dupire_europeans_720_payoffs_2^22_mrg32k3a_paths_x_12_steps_aad_multi_1081x720", "AADriskMulti", 1 << 12),
    5: (1 << 22, False, 1.2e4, "displaced_10_assets_autocall_2^22_mrg32k3a_paths_x_12_periods_aad_107_risks", "AADriskOne", 1 << 17),
}


def _call_config(api, k, model, n, sobol):
    if k == 4:
        return api.aad_risk_multi(model, "bench_prd", n, sobol=sobol)
    return api.aad_risk_one(model, "bench_prd", n, sobol=sobol)


def run_config(args):
    """One of the secondary configs through the host API.  N > 1: every rank calls the entry point with the full path
    count; the library runs the rank's shard and sums over the ranks inside its kernels (IPC communicator)."""
    import torch
    import torch.distributed as dist
    from compfinance_b200 import capi
    from compfinance_b200.api import CompFinance
    k = args.config
    n_paths, sobol, flops, workload, entry, cpu_sample = CONFIGS[k]
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cf = CompFinance(device=local_rank)
    eng = capi.Engine()
    model = _put_config(cf, k)
    if world > 1:
        from compfinance_b200.dist import connect_comm
        n_par = cf.num_params(model)
        n_pay = cf.num_payoffs("bench_prd")
        connect_comm(eng, max(4096, n_pay * (n_par + 600)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = None
    for _ in range(args.warmup):
        res = _call_config(cf, k, model, n_paths, sobol)
    torch.cuda.synchronize()
    launches0 = eng.lib.cf_launch_count()
    sampler.mark()
    if world > 1:
        dist.barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_tot, kms = 0.0, []
    for _ in range(args.steps):
        flush.fill_(1)                                           # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = _call_config(cf, k, model, n_paths, sobol)
        t_tot += time.perf_counter() - t0
        kms.append(eng.lib.cf_last_run_kernel_ms())
    tt = torch.tensor([t_tot], dtype=torch.float64, device="cuda")
    km = torch.tensor([float(np.mean(kms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    launches = eng.lib.cf_launch_count() - launches0
    if rank == 0:
        e2e_value = n_paths * args.steps / float(tt.item())
        kernel_ms = float(km.item())
        value = n_paths / (kernel_ms * 1e-3)                     # device time of the path kernels (max over ranks)
        fp64_peak = eng.fp64_peak_tflops()
        achieved = value * flops / 1e12 / world                  # per GPU: the peak is one GPU's
        cpu = None
        if not args.no_cpu_baseline:
            try:
                from oracle import refapi
                ref = refapi.get()
                threads = ref.start_pool(-1) + 1
                rmodel = _put_config(ref, k)
                _call_config(ref, k, rmodel, min(cpu_sample, 1 << 12), sobol)
                t0 = time.perf_counter()
                _call_config(ref, k, rmodel, cpu_sample, sobol)
                secs = time.perf_counter() - t0
                cpu = {"value": cpu_sample / secs, "unit": "paths/s", "cores": threads, "kind": "reference",
                       "sample": f"{cpu_sample} paths of the same workload through {entry}"}
            except (OSError, FileNotFoundError) as ex:
                cpu = {"value": None, "unit": "paths/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        values = np.asarray(res[0]).ravel()
        line = {
            "metric": "paths_per_sec_incl_full_aad_risk", "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": k, "paths": n_paths, "rng": "sobol" if sobol else "mrg32k3a", "entry_point": entry,
                       "parallelism": f"paths sharded over {world} GPU(s)" + (", rank sum over peer memory inside the reduction kernels" if world > 1 else ""),
                       "l2": "flushed between timed iterations (256 MB write)", "first_payoff_value": float(values[0])},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": 8 * cf.num_payoffs("bench_prd"),
                    "d2h_bytes_per_step": 8 * int(sum(np.asarray(r).size for r in res if not np.isscalar(r))), "steps": args.steps,
                    "api": f"{entry} (libcf_host.so)", "regime": "resident session"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": None, "kernel_ms": kernel_ms, "algorithmic_flops_per_path": flops,
                         "peak_source": "measured here: scalar DFMA microbenchmark cf_measure_fp64_peak",
                         "note": "value and kernel_ms are the device time of the path kernels of the call (CUDA events on the launch stream, max over ranks); achieved is per GPU; e2e is the wall time of the call"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE config: 3 (default) is the headline; the others are the parity configs at their full sizes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config != 3:
        args.steps = min(args.steps, 20)
        run_config(args)
        return

    import torch
    import torch.distributed as dist
    from compfinance_b200 import capi
    from compfinance_b200.api import CompFinance

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- set-up through the reference-facing host API; the resident plan comes from the same tables
    cf = CompFinance(device=local_rank)
    eng = capi.Engine()
    spots, times, vols = config3()
    cf.put_dupire(100.0, spots, times, vols, 0.25, "bench_dupire")
    cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "bench_uoc")
    d = cf.describe("bench_dupire", "bench_uoc", aad=True)
    mdl = eng.dupire_model(100.0, d["tab_b"], d["tab_a"], d["is_event"], d["n_events"], time_map=d["time_map"])
    prd = eng.uoc(d["strike"], d["barrier"], d["smooth"], d["n_events"])
    rng = eng.rng("sobol")
    plan = C.c_void_p()
    eng._chk(eng.lib.cf_plan_create(C.byref(mdl), C.byref(prd), C.byref(rng), C.byref(plan)))
    n_out = eng.lib.cf_plan_out_size(plan, 1)
    d_out = torch.zeros(n_out, dtype=torch.float64, device="cuda")
    wv = (C.c_double * 2)(1.0, 0.0)

    # strong scaling: disjoint skip-ahead blocks of the 2^20 paths
    from compfinance_b200.dist import shard_range
    first, count = shard_range(N_PATHS, rank, world)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    # The sum over ranks is part of the final reduction kernel (peer memory over NVLink, cf_plan_set_peers); it is
    # checked once against one NCCL all-reduce of the per-rank results, which is also the fallback.
    fused = False
    if world > 1:
        # reference result: per-rank launches, one NCCL all-reduce
        eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, first, count, d_out.data_ptr(), C.c_void_p(stream.cuda_stream)))
        dist.all_reduce(d_out)
        want = d_out.clone()
        ok_local = 1.0
        try:
            from compfinance_b200.dist import connect_comm
            connect_comm(eng, max(n_out, 4096))
            for _ in range(2):
                eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, first, count, d_out.data_ptr(), C.c_void_p(stream.cuda_stream)))
            torch.cuda.synchronize()
            ok_local = 1.0 if torch.allclose(d_out, want, rtol=1e-11, atol=1e-9) else 0.0
        except Exception as ex:                                   # the blocks cannot be shared: NCCL per step
            print(f"bench.py: rank {rank}: peer-memory reduction unavailable ({type(ex).__name__}: {ex})", file=sys.stderr)
            ok_local = 0.0
        okf = torch.tensor([ok_local], device="cuda")
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        fused = bool(okf.item() > 0.5)
        if not fused:
            eng.lib.cf_comm_enable(0)
            if rank == 0:
                print("bench.py: falling back to one NCCL all-reduce per step", file=sys.stderr)

    def step(first_path=first, n_paths=count):
        eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, first_path, n_paths, d_out.data_ptr(), C.c_void_p(stream.cuda_stream)))
        if world > 1 and not fused:
            dist.all_reduce(d_out)

    align = torch.zeros(1, device="cuda")

    def timed(n_steps, first_path, n_paths):
        """n_steps passes, each bracketed by CUDA events on the launching stream; returns ms per step, max over ranks."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(n_steps):
            flush.fill_(1)                                       # L2 flush between timed iterations (untimed)
            if world > 1:
                dist.all_reduce(align)                           # untimed: the ranks' streams leave the flush together, so that
                                                                 # a step is timed from a common start (max over ranks below)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(first_path, n_paths)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps

    sampler = ClockSampler(local_rank)
    sampler.start()                                              # nvidia-smi needs ~0.1 s to deliver its first sample
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    eng._chk(eng.lib.cf_plan_kernel_ms(plan, None, None))       # drop warm-up kernel timings
    launches0 = eng.lib.cf_launch_count()
    sampler.mark()                                               # keep the samples from here on: timed region, (weak pass,) e2e loop
    ms_per_step = timed(args.steps, first, count)
    launches = eng.lib.cf_launch_count() - launches0
    value = N_PATHS / (ms_per_step * 1e-3)
    kms, kn = C.c_double(), C.c_int()
    eng._chk(eng.lib.cf_plan_kernel_ms(plan, C.byref(kms), C.byref(kn)))
    kms_all = torch.tensor([kms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        gathered = [torch.zeros_like(kms_all) for _ in range(world)]
        dist.all_gather(gathered, kms_all)
        kms_all = torch.cat(gathered)
    kms_per_rank = [round(float(v), 5) for v in kms_all.cpu()]
    res = d_out.cpu().numpy()
    price, delta = res[2] / N_PATHS, res[3] / N_PATHS

    # ---- the same pass with 2^20 paths PER GPU (weak scaling), reported beside the strong-scaling headline
    weak = None
    if world > 1:
        wfirst, wcount = shard_range(N_PATHS * world, rank, world)
        for _ in range(args.warmup):
            step(wfirst, wcount)
        wms = timed(args.steps, wfirst, wcount)
        eng._chk(eng.lib.cf_plan_kernel_ms(plan, None, None))
        wres = d_out.cpu().numpy()
        weak = {"paths_per_gpu": N_PATHS, "paths": N_PATHS * world, "ms_per_step": wms, "value": N_PATHS * world / (wms * 1e-3),
                "unit": "paths/s", "price": wres[2] / (N_PATHS * world)}

    # ---- e2e: the call a user makes (dupireAADRisk through the host API, host buffers in and out)
    e2e_steps = max(3, args.steps)                              # as many end-to-end calls as timed steps: the closing barrier is amortised alike
    # host -> device per call: the notionals (kernel parameters); the model's tables went up when the session of this
    # (model, product, RNG) was built by the first call after putDupire / putBarrier and stay resident (cf_base.h).
    # e2e["first_call"] times the other regime: the session dropped before every call (clone, init() on the host tape,
    # device images, 62 KB of tables uploaded, plan created -- every step).
    h2d_tables = int(d["tab_a"].nbytes + d["tab_b"].nbytes + d["is_event"].nbytes + 4 * 156 * 8 + 32 * 156 * 4 + 2 * 8)
    h2d = 16
    d2h = int(n_out * 8)

    def e2e_step():
        # the call a user makes, on every rank: with the communicator connected the library runs this rank's shard of the
        # 2^20 paths and the rank sum is part of its kernels; every rank returns the full result
        if world == 1 or fused:
            return cf.dupire_aad_risk("bench_dupire", "bench_uoc", [1.0, 0.0], 30, 36, N_PATHS)
        r = eng.run_aad(mdl, prd, rng, first, count, [1.0, 0.0])          # C ABI, host buffers
        v = torch.from_numpy(np.concatenate([r["payoff_sums"], [r["agg_sum"]], r["table_adj"]])).pin_memory().cuda(non_blocking=True)
        dist.all_reduce(v)
        return v.cpu().numpy()

    e2e_res = e2e_step()
    e2e_check = None
    if isinstance(e2e_res, tuple):                                # (value, delta, vega) of dupireAADRisk: the same numbers as the device leg
        e2e_check = {"price_rel_diff_vs_device_leg": abs(e2e_res[0] / price - 1), "delta_rel_diff_vs_device_leg": abs(e2e_res[1] / delta - 1)}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N_PATHS * e2e_steps / float(te.item())
    # the same with the resident session dropped before every call
    first_call_value = None
    if world == 1 or fused:
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            cf.lib.cfx_drop_sessions()
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tf = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        first_call_value = N_PATHS * e2e_steps / float(tf.item())
    clocks = sampler.stop()          # sampled under load: warm-up, timed region and the e2e loop

    if rank == 0:
        # the roofline denominator, measured here with its own clock record: the DFMA microbenchmark is repeated for ~0.2 s
        peak_sampler = ClockSampler(local_rank)
        peak_sampler.start()
        fp64_peak, t_end = 0.0, time.perf_counter() + 0.2
        while time.perf_counter() < t_end:
            fp64_peak = max(fp64_peak, eng.fp64_peak_tflops())
        peak_clocks = peak_sampler.stop()
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        nominal_peak = 148 * 64 * 2 * sm_max * 1e6 / 1e12           # 148 SMs x 64 DFMA / clk x 2 flop at the maximum SM clock
        achieved = (count / (kms.value * 1e-3)) * FLOPS_PER_PATH / 1e12 if kms.value > 0 else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "latest_kernel.json")))
            traffic = prof.get("dram_bytes_per_launch")       # ncu capture of one 2^20-path launch pair
            if traffic and world > 1:
                traffic = traffic * count / N_PATHS            # a rank's shard: history traffic is proportional to its paths
        except OSError:
            pass
        roofline = {
            "bound": "fp64", "kernel": ("cf::dupire_forward4_kernel<UOC, AAD, Sobol, 2 paths / thread, 28 warps> + cf::dupire_reverse_kernel<UOC>" if count > 148 * 1536 else "cf::dupire_forward4_kernel<UOC, AAD, Sobol, 1 path / thread, 8-step chunks> + cf::dupire_reverse_span_kernel<UOC, 5> (programmatic dependent launch)") + " (one CUDA-event bracket around the pair)", "achieved": achieved, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": achieved / fp64_peak if achieved else None, "traffic": traffic,
            "kernel_ms": kms.value, "kernel_ms_per_rank": kms_per_rank, "kernel_launches_timed": kn.value,
            "peak_source": "measured here: scalar DFMA microbenchmark cf_measure_fp64_peak, best of ~0.2 s of launches (MEASURED_PEAKS.json has no fp64 entry)",
            "peak_clocks": peak_clocks, "peak_nominal": nominal_peak,
            "peak_nominal_source": "148 SMs x 64 DFMA/clk x 2 flop x max SM clock", "frac_of_nominal": achieved / nominal_peak if achieved else None,
            "traffic_source": "from_profile (profiles/latest_kernel.json: ncu --set full capture of one 2^20-path launch pair, not measured in this run)",
            "algorithmic_flops_per_path": FLOPS_PER_PATH,
            "hbm": {"achieved_gbs": (traffic / (kms.value * 1e-3) / 1e9) if traffic and kms.value > 0 else None,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json" if peaks else None},
        }
        cpu = None
        if not args.no_cpu_baseline:
            try:
                pps, threads, rv, rd, secs = cpu_reference_run(1 << 20, repeats=2)
                cpu = {"value": pps, "unit": "paths/s", "cores": threads, "kind": "reference",
                       "sample": "2^20 paths x 156 steps (the full workload), best of 2",
                       "price_rel_diff_vs_gpu": abs(price / rv - 1), "delta_rel_diff_vs_gpu": abs(delta / rd - 1)}
            except (OSError, FileNotFoundError) as ex:
                cpu = {"value": None, "unit": "paths/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        line = {
            "metric": "paths_per_sec_incl_full_aad_risk", "value": value, "unit": "paths/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "paths": N_PATHS, "steps_per_path": 156, "surface": "30x36", "rng": "sobol",
                       "risks": 1081, "parallelism": (f"paths sharded over {world} GPU(s), sum of {n_out} doubles over ranks inside the reduction kernel (peer memory over NVLink)" if fused else f"paths sharded over {world} GPU(s), one NCCL all-reduce of {n_out} doubles"),
                       "l2": "flushed between timed iterations (256 MB write)" + ("; ranks aligned by an untimed all-reduce after each flush" if world > 1 else ""), "price": price, "delta": delta},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": h2d,
                                      "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                                      "api": "dupireAADRisk (libcf_host.so)" if (world == 1 or fused) else "cf_run_aad (C ABI, host buffers) + NCCL all-reduce",
                                      "value_check": e2e_check,
                                      "regime": "resident session: tables uploaded by the first call after putDupire, reused since (h2d per step = the notionals)",
                                      "first_call": {"value": first_call_value, "unit": "paths/s", "h2d_bytes_per_step": h2d_tables + 16,
                                                     "regime": "session dropped before every call: clone + init() on the host tape + table upload + plan, every step"}},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if weak:
            line["weak_scaling"] = weak
        print(json.dumps(line))
    eng.lib.cf_plan_destroy(plan)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
