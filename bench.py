#!/usr/bin/env python
"""bench.py — plskern fit on B200 (BASELINE.json metric: fit seconds & FP64 TFLOP/s at
n=1e6, p=500, q=10, nlv=25, Float64, `rand` data), N = 1/2/4/8 GPUs of one node.

  python bench.py --gpus 1 --steps K --warmup W            # our arm, one process per GPU under torchrun
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference algorithm on host cores

A step = one `plskern` fit (pivot, K1 Gram, [all-reduce], K3/K4 solve, K5 scores, weights) of the
configured shape.  Weak scaling: every GPU holds n_per_gpu = 1e6 rows, the global matrix has N*1e6.
`value` = algorithmic F_fit(n_global) / max-over-ranks device time, inputs resident in HBM.
`e2e`   = same metric through the host-pointer C ABI (N=1: jcb200_plskern_fit on pinned host arrays;
N>1: pinned-host -> device copies + the sharded fit + device -> host of T per rank).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[1]
N_PER_GPU, P, Q, NLV = 1_000_000, 500, 10, 25
README_PLSKERN_SECONDS = 8.100469      # /root/reference/README.md:91 (i9-10885H laptop), BASELINE.md §1
CPU_SAMPLE_ROWS = int(os.environ.get("JCB_BENCH_CPU_ROWS", 250_000))    # bounded sample for the CPU arms


def f_gram(n, p, q):
    return n * p * (p + 1) + 2.0 * n * p * q


def f_fit(n, p, q, nlv):
    """upper triangle of X'DX once + X'DY + scores + the (negligible) LV loop."""
    f_lv = nlv * (2.0 * p * p + 2.0 * p * q * q + 6.0 * p * q) + 2.0 * p * nlv * nlv
    return f_gram(n, p, q) + 2.0 * n * p * nlv + f_lv


def load_peaks():
    out = {}
    try:
        out.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except Exception:
        pass
    try:
        out["fp64"] = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
    except Exception:
        out["fp64"] = None
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, uuid):
        self.uuid, self.proc, self.lines = uuid, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        clocks, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    clocks.append(float(f[0]))
                    power.append(float(f[2]))
                    for nm, v in zip(names, f[3:7]):
                        if v == "Active":
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": mx,
                "power_w_max": max(power) if power else None, "samples": len(clocks),
                "reasons": sorted(reasons)}


def cpu_fit_seconds(n_rows, reps, warmup, threads):
    """The NumPy oracle (restatement of the reference algorithm) timed on the host cores."""
    import numpy as np
    from threadpoolctl import threadpool_limits
    import oracle
    from oracle import synth
    X = synth.synth_matrix(synth.SEED_X, n_rows, P)
    Y = synth.synth_matrix(synth.SEED_Y, n_rows, Q)
    times = []
    with threadpool_limits(limits=threads):
        for i in range(warmup + reps):
            Xc, Yc = X.copy(order="F"), Y.copy(order="F")
            t0 = time.perf_counter()
            oracle.plskern_bang(Xc, Yc, None, nlv=NLV)     # plskern!: no input copy inside the timing
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, np.__version__


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Julia is not installed on this
    image, so this is the oracle port (kind = "port") with all host threads, on a bounded row sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_s = CPU_SAMPLE_ROWS
    times, npver = cpu_fit_seconds(n_s, args.steps, args.warmup, threads)
    sec = sum(times) / len(times)
    val = f_fit(n_s, P, Q, NLV) / sec * 1e-12
    sample = (f"NumPy {npver} restatement of plskern! (oracle/plskern_ref.py), first {n_s} rows of the "
              f"workload (p={P}, q={Q}, nlv={NLV}), {threads} BLAS threads; TFLOP/s is size-normalised")
    line = {
        "impl": "reference", "metric": "plskern_fit_fp64_tflops", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"plskern fit n={n_s} (bounded sample of n=1e6 per GPU) p={P} q={Q} "
                               f"nlv={NLV} Float64 uniform weights scal=false", "cpu_only": True},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": "port",
                         "sample": sample, "fit_seconds_sample": sec,
                         "fit_seconds_scaled_to_1e6_rows": sec * (N_PER_GPU / n_s)},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (e.g. NCCL's version banner) must not reach stdout: the contract is ONE JSON line.
    fd 1 is pointed at stderr for the whole run; emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "jchemo.jl_b200", "libjchemo_b200.so")):
        ge.build()
    import jchemo_b200 as jc
    from jchemo_b200 import _lib, device as dev, sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
    dev.init(local_rank)
    dev.use_current_stream()
    lib = jc.lib()

    n_loc = args.rows_per_gpu
    n_glob = n_loc * world
    row0 = rank * n_loc
    K, W = args.steps, max(args.warmup, 3)

    # ---- synthetic inputs, generated on the device (identical bits to oracle/synth.py)
    X = dev.colmajor_empty(n_loc, P, device)
    Y = dev.colmajor_empty(n_loc, Q, device)
    dev.fill_uniform(X, n_loc, 1, row0, n_glob)
    dev.fill_uniform(Y, n_loc, 2, row0, n_glob)
    model = dev.DeviceModel(n_loc, P, Q, NLV, device)
    pivot = torch.empty(P + Q + 1, dtype=torch.float64, device=device)
    packed = torch.empty(dev.packed_len(P, Q), dtype=torch.float64, device=device)

    def step():
        sharded.fit_sharded(X, Y, None, n_loc, model, scal=False, pivot=pivot, packed=packed)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    uuid = str(torch.cuda.get_device_properties(device).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.jcb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = lib.jcb200_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    gram_ms = _lib.gram_timings(K)
    phases = dev.sync_timings()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = f_fit(n_glob, P, Q, NLV) / (ms_step * 1e-3) * 1e-12

    # ---- end to end through the host-pointer path (pinned host buffers, copies inside the timing)
    Ke = max(1, min(args.e2e_steps, K))
    numa_node = None
    if world > 1 and os.environ.get("JCB_NUMA_BIND", "1") != "0":
        numa_node = sharded.bind_host_to_gpu_numa(local_rank)     # page-locked buffers local to the GPU's root
    hX = torch.empty((P, n_loc), dtype=torch.float64).pin_memory()
    hY = torch.empty((Q, n_loc), dtype=torch.float64).pin_memory()
    hX.copy_(X[:, :n_loc])
    hY.copy_(Y[:, :n_loc])
    h2d = 8 * (n_loc * P + n_loc * Q)
    d2h = 8 * (n_loc * NLV + 3 * P * NLV + Q * NLV + NLV + 2 * P + 2 * Q + n_loc)
    if world == 1:
        Xh, Yh = hX.numpy().T, hY.numpy().T          # column-major [n, p] views of the pinned buffers
        assert Xh.flags.f_contiguous and Yh.flags.f_contiguous
        dev.use_own_stream()                         # the C ABI call times itself on its own stream
        for _ in range(3):                           # warm-up: device buffers and the pinned output pool
            fm = jc.plskern(Xh, Yh, nlv=NLV)         # (bound like in the timed loop: two blocks alternate)
        t0 = time.perf_counter()
        for _ in range(Ke):
            fm = jc.plskern(Xh, Yh, nlv=NLV)
        e2e_s = (time.perf_counter() - t0) / Ke
        e2e_phases = jc.last_timings()
        dev.use_current_stream()
        e2e_how = "jcb200_plskern_fit (C ABI, host pointers) on pinned numpy arrays, wall clock"
    else:
        hT = torch.empty((NLV, n_loc), dtype=torch.float64).pin_memory()
        small = [torch.empty_like(x, device="cpu").pin_memory() for x in
                 (model.P, model.R, model.W, model.C, model.TT, model.xmeans, model.xscales,
                  model.ymeans, model.yscales, model.weights)]

        def e2e_step():
            # this rank's rows streamed from page-locked host memory under K1, scores copied back under K5
            sharded.fit_sharded_from_host(hX, hY, None, X, Y, None, n_loc, model, scal=False, pivot=pivot,
                                          packed=packed, hT=hT)
            for h, d in zip(small, (model.P, model.R, model.W, model.C, model.TT, model.xmeans,
                                    model.xscales, model.ymeans, model.yscales, model.weights)):
                h.copy_(d, non_blocking=True)
        e2e_step()
        barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(Ke):
            e2e_step()
        ee1.record()
        barrier()
        te = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device=device)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item()) / Ke * 1e-3
        e2e_phases = None
        e2e_how = ("per rank: sharded.fit_sharded_from_host (row chunks from pinned host memory under K1, one packed-Gram "
                   "all-reduce, scores copied back in row blocks under K5) + device -> host of the model")
    e2e_val = f_fit(n_glob, P, Q, NLV) / e2e_s * 1e-12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fp64 = peaks.get("fp64") or {}
    peak_tf = fp64.get("dmma_tflops_burst")
    gram_avg = sum(gram_ms) / len(gram_ms) if gram_ms else None
    achieved = f_gram(n_loc, P, Q) / (gram_avg * 1e-3) * 1e-12 if gram_avg else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic_r01.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "kernel": "gram_kernel<false> (K1: fused centre + DMMA SYRK/GEMM)", "bound": "tensor",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (achieved / peak_tf) if (achieved and peak_tf) else None, "traffic": traffic,
        "peak_source": "measured FP64 DMMA.8x8x4 register-resident loop on this pool's B200 "
                       "(profiles/fp64_peak_r01.json; MEASURED_PEAKS.json holds no FP64 figure)",
        "algorithmic_flops_per_launch": f_gram(n_loc, P, Q), "avg_launch_ms": gram_avg,
        "launches_timed": len(gram_ms), "share_of_step": (gram_avg / ms_step) if gram_avg else None,
    }
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:          # the CPU leg is timed at N=1 only (rank 0)
        threads = os.cpu_count() or 1
        times, npver = cpu_fit_seconds(CPU_SAMPLE_ROWS, 3, 1, threads)
        sec = sum(times) / len(times)
        cpu_baseline = {
            "value": f_fit(CPU_SAMPLE_ROWS, P, Q, NLV) / sec * 1e-12, "unit": "TFLOP/s",
            "cores": threads, "kind": "port",
            "sample": f"NumPy {npver} restatement of plskern! on the first {CPU_SAMPLE_ROWS} rows "
                      f"(p={P}, q={Q}, nlv={NLV}), {threads} BLAS threads, mean of 3 after 1 warm-up",
            "fit_seconds_sample": sec, "fit_seconds_scaled_to_1e6_rows": sec * (N_PER_GPU / CPU_SAMPLE_ROWS),
        }
    base_tf = f_fit(N_PER_GPU, P, Q, NLV) / README_PLSKERN_SECONDS * 1e-12
    line = {
        "metric": "plskern_fit_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_step, "fit_seconds": ms_step * 1e-3,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": (value / base_tf) if world == 1 and n_loc == N_PER_GPU else None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"plskern fit n={n_glob} ({n_loc} rows per GPU) p={P} q={Q} nlv={NLV} "
                               "Float64 uniform weights scal=false (BASELINE.json configs[1])",
                   "parallelism": f"rows sharded over {world} GPU(s); one packed-Gram all-reduce",
                   "l2": "inputs (4.08 GB per GPU) are larger than L2; no flush needed",
                   "vs_baseline_def": "value / (F_fit / 8.100469 s README plskern time on an i9-10885H)"},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_val, "unit": "TFLOP/s", "fit_seconds": e2e_s,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke, "how": e2e_how,
                "phases_ms": e2e_phases, "numa_node_rank0": numa_node,
                "vs_readme_seconds": (README_PLSKERN_SECONDS / e2e_s) if world == 1 else None},
        "gpu_launches": int(launches), "phases_ms_last_step": phases, "clocks": clocks,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
