#!/usr/bin/env python
"""bench.py — KMeans.fit() Lloyd-loop throughput on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (sm_100a kernels via the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm: the oracle port on host cores

A "step" is ONE Lloyd iteration over this rank's resident partition: fused assign+partial-sum pass over X,
fixed-order partial reduce, NCCL allreduce of the [k*d sums | k counts | cost] buffer (N>1), finalize.
Workload at every N: BASELINE.json configs[1] PER GPU (k=64, n=10M rows/GPU, d=128, float32; weak scaling).
Prints ONE JSON line (rank 0).  Contract details: DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n_per_gpu, d, k)
    "cfg2": (10_000_000, 128, 64),
    "cfg3": (12_500_000, 256, 256),
    "small": (1_000_000, 128, 64),
}
METRIC = "kmeans_fit_samples_per_sec"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--n-per-gpu", type=int, default=0, help="override rows per GPU")
    ap.add_argument("--kernel-path", default="auto", choices=["auto", "generic", "tcgen05"])
    ap.add_argument("--init", default="first_k", choices=["first_k", "near_true"],
                    help="initial centres of the timed Lloyd loop: the first k rows of rank 0, or the generating "
                         "centres + 0.25 sigma noise (what a k-means|| start looks like on separated blobs)")
    ap.add_argument("--probe", type=int, default=0, help="diagnostic builds only (B2K_LIB=libb2kmeans_probe.so)")
    ap.add_argument("--e2e-iters", type=int, default=20, help="maxIter of the end-to-end fit (Spark default 20)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(workload_rows: int, d: int):
    """dram bytes per launch of the fused kernel from the committed ncu summary (profiles/), scaled to this n."""
    p = os.path.join(ROOT, "profiles", "fused_kernel_ncu_summary.json")
    if not os.path.exists(p):
        return None
    try:
        j = json.load(open(p))
        per_row = float(j["dram_bytes_per_row"])
        if int(j.get("d", d)) != d:
            return None
        return per_row * workload_rows
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (r & bit):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


_CPU_SAMPLE = {}


def cpu_sample(n_rows: int, d: int, k: int, seed: int = 1234):
    """Host blobs of the benchmark's shape (k centres ~ U(-10,10)^d, unit-variance noise), generated once per process."""
    import numpy as np

    key = (n_rows, d, k, seed)
    if key not in _CPU_SAMPLE:
        rng = np.random.default_rng(seed)
        centers = rng.uniform(-10.0, 10.0, size=(k, d)).astype(np.float32)
        X = rng.standard_normal(size=(n_rows, d), dtype=np.float32)
        X += centers[rng.integers(0, k, size=n_rows)]
        _CPU_SAMPLE.clear()
        _CPU_SAMPLE[key] = np.ascontiguousarray(X)
    return _CPU_SAMPLE[key]


def cpu_reference_run(n_rows: int, d: int, k: int, iters: int, seed: int = 1234):
    """Times the oracle's C/OpenMP Lloyd port (oracle/kmeans_oracle.c) on a bounded sample — CPU baseline only."""
    from oracle import c_oracle

    X = cpu_sample(n_rows, d, k, seed)
    C0 = X[:k].copy()
    c_oracle.lloyd(X[: min(n_rows, 2000)], C0, 1, -1.0, want_labels=False)  # warm the library
    t0 = time.perf_counter()
    out = c_oracle.lloyd(X, C0, iters, -1.0, want_labels=False)
    dt = time.perf_counter() - t0
    assert out["n_iter"] == iters
    return n_rows * iters / dt, dt, c_oracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, k = CONFIGS[args.config]
    rows = args.cpu_sample_rows
    cpu_reference_run(rows, d, k, max(1, args.warmup))
    val, dt, threads = cpu_reference_run(rows, d, k, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: KMeans k={k} d={d} float32 blobs; each step = one Lloyd "
                               f"iteration over a bounded {rows}-row sample on host cores",
                   "k": k, "d": d, "sample_rows": rows},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{rows} rows x {args.steps} Lloyd iterations, oracle/kmeans_oracle.c (OpenMP)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "pyspark.ml.clustering.KMeans (BASELINE configs[0]) cannot run here: no pyspark/JVM in the image; "
                "the reference's GPU arithmetic (cuML) is absent too, so the CPU arm is the oracle port",
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    from spark_rapids_ml_b200 import _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the KMeans path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # type: ignore

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_local, d, k = CONFIGS[args.config]
    if args.n_per_gpu:
        n_local = args.n_per_gpu
    n_total = n_local * world

    ctx = _native.Context(local_rank)
    ctx.set_option("kernel_path", {"auto": 0, "generic": 1, "tcgen05": 2}[args.kernel_path])
    if args.probe:
        ctx.set_option("probe", args.probe)
    if world > 1:
        uid = torch.zeros(_native.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(_native.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))

    # ---- synthetic blobs on the device: centers U(-10,10)^d shared by all ranks ----
    g = torch.Generator(device=dev).manual_seed(42)
    centers_true = torch.rand((k, d), generator=g, device=dev) * 20.0 - 10.0
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.empty((n_local, d), dtype=torch.float32, device=dev)
    chunk = 1_000_000
    for s in range(0, n_local, chunk):
        e = min(n_local, s + chunk)
        z = torch.randint(0, k, (e - s,), generator=g, device=dev)
        X[s:e] = centers_true[z] + torch.randn((e - s, d), generator=g, device=dev)
    if args.init == "near_true":
        g0 = torch.Generator(device=dev).manual_seed(7)
        C0 = (centers_true + 0.25 * torch.randn((k, d), generator=g0, device=dev)).contiguous()
    else:
        C0 = X[:k].clone()
    if world > 1:
        dist.broadcast(C0, 0)  # deterministic "array" init = first k rows of rank 0

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up ----
    C = C0.clone()
    ctx.kmeans_lloyd(X, C, max(args.warmup, 3), -1.0)

    # ---- timed region: exactly K Lloyd iterations (tol < 0 never converges), device-timed ----
    C = C0.clone()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.set_option("time_kernels", 1)   # CUDA events around every fused launch of THIS timed loop (roofline below)
    ctx.set_option("collect_recheck", 1)
    e0.record()
    n_iter, shift = ctx.kmeans_lloyd(X, C, args.steps, -1.0)
    e1.record()
    torch.cuda.synchronize(dev)
    ctx.set_option("time_kernels", 0)
    clocks = sampler.stop()
    barrier()
    assert n_iter == args.steps, (n_iter, args.steps)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    st = ctx.stats()
    launches = int(st["kernel_launches"])
    path = {1: "generic", 2: "tcgen05"}.get(st["last_path"], "?")
    value = n_total * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel: live CUDA-event timing of each fused launch ----
    roofline = None
    peak, peak_src = measured_peaks()
    st2 = st   # the per-launch event times of the timed loop itself
    if st2["last_fused_ms"] > 0:
        alg_bytes = 4.0 * n_local * d  # X read once; partial flush 148*(k*d+k)*4 B is < 0.1 %
        ach = alg_bytes / (st2["last_fused_ms"] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_fused_assign_update", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": ncu_traffic_per_launch(n_local, d), "peak_source": peak_src,
                    "kernel_ms": st2["last_fused_ms"], "loop_ms_per_iter": st2["last_loop_ms"] / max(1, st2["last_n_iter"]),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "tensor_flops_per_launch": 2.0 * n_local * d * k,
                    "tensor_tflops_1x": 2.0 * n_local * d * k / (st2["last_fused_ms"] / 1e3) / 1e12,
                    "recheck_rows_per_iter": st2["recheck_rows"] / max(1, st2["last_n_iter"]),
                    "recheck_candidates_per_iter": st2["recheck_candidates"] / max(1, st2["last_n_iter"])}
    else:
        # generic path: time one assign+update iteration as a whole
        roofline = {"bound": "hbm", "kernel": "generic assign+update (2 passes over X)", "achieved":
                    4.0 * n_local * d / (ms / args.steps / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": 4.0 * n_local * d / (ms / args.steps / 1e3) / 1e9 / peak, "traffic": None,
                    "peak_source": peak_src}

    # ---- end to end through the C ABI with HOST buffers: pinned host X -> ingest (H2D) -> fit -> centers D2H ----
    e2e = None
    if not args.no_e2e:
        try:
            Xh = torch.empty((n_local, d), dtype=torch.float32, pin_memory=True)
            Xh.copy_(X)
            C0h = C0.cpu().pin_memory()
            torch.cuda.synchronize(dev)
            Xd = torch.empty_like(X)
            del X
            times = []
            iters_done = 0
            for rep in range(args.e2e_steps + 1):  # first rep is warm-up
                barrier()
                t0 = torch.cuda.Event(enable_timing=True)
                t1 = torch.cuda.Event(enable_timing=True)
                t0.record()
                ctx.ingest_pinned_tensor(Xd, 0, Xh)                       # H2D of this step's inputs
                C0d = C0h.to(dev, non_blocking=True)
                out = ctx.kmeans_fit(Xd, k, init=C0d, max_iter=args.e2e_iters, tol=1e-30, compute_inertia=False)
                res = out["cluster_centers_"].cpu()                        # D2H of the step's result
                t1.record()
                torch.cuda.synchronize(dev)
                barrier()
                if rep > 0:
                    times.append(t0.elapsed_time(t1))
                    iters_done = out["n_iter_"]
            t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
            e2e = {"value": n_total * iters_done / (e2e_ms / 1e3), "unit": UNIT,
                   "h2d_bytes_per_step": int(n_local * d * 4 + k * d * 4), "d2h_bytes_per_step": int(k * d * 4 + 64),
                   "ms_per_fit": e2e_ms, "iterations_per_fit": iters_done,
                   "what": "b2k_ingest_append(pinned host X) + b2k_kmeans_fit(init=array, maxIter=%d) + centers D2H, per fit"
                           % args.e2e_iters}
            del Xh, Xd
        except Exception as ex:  # pinned allocation can fail on small hosts: report, do not fake
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    # ---- CPU baseline (oracle port) on rank 0 at N=1, bounded sample ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows, cpu_iters = args.cpu_sample_rows, 20
        val, dt, threads = cpu_reference_run(rows, d, k, cpu_iters)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{rows} rows x {cpu_iters} Lloyd iterations of the same blobs shape (k={k}, d={d}), "
                                  f"oracle/kmeans_oracle.c OpenMP fp64, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: KMeans Lloyd iteration, k={k}, n={n_local}/GPU x {world} GPU, d={d}, "
                                   "float32 blobs resident in HBM; fixed 'array' init; tol<0 so every step does full work",
                       "k": k, "d": d, "n_per_gpu": n_local, "n_total": n_total, "kernel_path": path,
                       "l2": f"inputs ({n_local * d * 4 / 1e9:.2f} GB/GPU) are larger than the 126 MB L2: no flush needed",
                       "parallelism": f"dp{world} (rows sharded; one f64 allreduce of k*d+k+1 values per step)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches, "final_shift": shift,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
