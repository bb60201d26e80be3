#!/usr/bin/env python
"""bench.py — plskern fit on B200 (BASELINE.json metric: fit seconds & FP64 TFLOP/s at n=1e6, p=500, q=10,
nlv=25, Float64, `rand` data) at N = 1/2/4/8 GPUs of one node.

  python bench.py --gpus 1 --steps K --warmup W            # our arm, one process per GPU under torchrun
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference algorithm on host cores

A step = one `plskern` fit (pivot, K1 Gram, [exchange], K3/K4 solve, K5 scores, weights).
Headline = STRONG scaling of the metric's own configuration: the global matrix has n = 1e6 rows at every N,
rank r holds rows [r n/N, (r+1) n/N) in HBM.  `value` = algorithmic F_fit(1e6) / max-over-ranks device time.
The one exchange of the path (pivot of rank 0 + sum of the packed partial Grams) runs through peer HBM
(jcb200_comm_*: CUDA-IPC windows over NVLink, no collective call); the same step with an NCCL all-reduce is
timed beside it.  Extra legs on the same line: `weak` (1e6 rows per GPU, round 1's number), `parity` (the c2_cut
golden fitted row-sharded over the N ranks vs tests/golden/c2_cut.npz), `extra_configs.c4` (BASELINE configs[3]:
n=1e7, p=2000, q=10, nlv=50 over N >= 2 GPUs, generated on the device).
`e2e` = the same metric through the host-pointer C ABI from host arrays, copies inside the timing: N=1
jcb200_plskern_fit on one GPU; N>1 the same call with the library bound to all N GPUs (jcb200_init_multi, rank 0
drives them while the other ranks idle) — the call a Julia drop-in user makes on an N-GPU box.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[1] (the configuration the metric is quoted on) and configs[3]
N_GLOBAL, P, Q, NLV = 1_000_000, 500, 10, 25
C4 = dict(n=10_000_000, p=2000, q=10, nlv=50)
README_PLSKERN_SECONDS = 8.100469      # /root/reference/README.md:91 (i9-10885H laptop), BASELINE.md §1
README_PLSKERN_BANG_SECONDS = 7.232234  # /root/reference/README.md:94
CPU_ROWS = int(os.environ.get("JCB_BENCH_CPU_ROWS", N_GLOBAL))      # the CPU legs run the FULL configuration


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
    for name in ("fp64_peak_r02.json", "fp64_peak_r01.json"):
        try:
            out["fp64"] = json.load(open(os.path.join(ROOT, "profiles", name)))
            out["fp64"]["file"] = "profiles/" + name
            break
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


# ------------------------------------------------------------------------------------------ CPU legs
def julia_reference_seconds(n_rows, reps, warmup):
    """The UNMODIFIED reference (src/utility.jl + src/plskern.jl included by oracle/julia_ref.jl) when both a
    `julia` binary and the reference tree exist on this machine; None otherwise (this image has no Julia, and
    /root/reference does not travel to the GPU box)."""
    jl = shutil.which("julia")
    ref = os.environ.get("JCHEMO_REFERENCE", "/root/reference")
    script = os.path.join(ROOT, "oracle", "julia_ref.jl")
    if not (jl and os.path.isdir(os.path.join(ref, "src")) and os.path.exists(script)):
        return None
    try:
        res = subprocess.run([jl, "--startup-file=no", script, ref, "time", str(n_rows), str(P), str(Q), str(NLV),
                              str(reps), str(warmup)], capture_output=True, text=True, timeout=1500)
        if res.returncode != 0:
            return None
        out = json.loads(res.stdout.strip().splitlines()[-1])
        return out
    except Exception:
        return None


def cpu_fit_seconds(n_rows, reps, warmup, threads, X=None, Y=None):
    """The NumPy oracle (restatement of the reference algorithm, plskern!) timed on the host cores."""
    import numpy as np
    from threadpoolctl import threadpool_limits
    import oracle
    from oracle import synth
    if X is None:
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


def cpu_baseline_record(reps, warmup, X=None, Y=None):
    threads = os.cpu_count() or 1
    n_s = CPU_ROWS
    jl = julia_reference_seconds(n_s, reps, warmup)
    if jl is not None:
        sec = jl["seconds_mean"]
        return {"value": f_fit(n_s, P, Q, NLV) / sec * 1e-12, "unit": "TFLOP/s", "cores": jl.get("blas_threads", threads),
                "kind": "reference", "fit_seconds": sec,
                "sample": f"unmodified Jchemo plskern! under Julia {jl.get('julia')} on n={n_s} p={P} q={Q} nlv={NLV}, "
                          f"BLAS threads {jl.get('blas_threads')}, mean of {reps} after {warmup} warm-up"}, sec
    times, npver = cpu_fit_seconds(n_s, reps, warmup, threads, X, Y)
    sec = sum(times) / len(times)
    full = "the full configuration" if n_s == N_GLOBAL else f"the first {n_s} rows (JCB_BENCH_CPU_ROWS)"
    return {"value": f_fit(n_s, P, Q, NLV) / sec * 1e-12, "unit": "TFLOP/s", "cores": threads, "kind": "port",
            "fit_seconds": sec, "fit_seconds_min_max": [min(times), max(times)],
            "sample": f"NumPy {npver} restatement of plskern! (oracle/plskern_ref.py) on {full}: n={n_s} p={P} q={Q} "
                      f"nlv={NLV}, {threads} BLAS threads, mean of {reps} fit(s) after {warmup} warm-up"}, sec


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores, SAME
    configuration as our arm (n = 1e6 rows).  Julia + the reference tree when both exist here, else the oracle
    port (kind = "port") with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rec, sec = cpu_baseline_record(args.steps, args.warmup)
    val = rec["value"]
    line = {
        "impl": "reference", "metric": "plskern_fit_fp64_tflops", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "fit_seconds": sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(CPU_ROWS, 1, cpu=True),
        "cpu_baseline": rec,
        "e2e": {"value": val, "unit": "TFLOP/s", "fit_seconds": sec, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n, world, cpu=False):
    cfg = {"workload": f"plskern fit n={n} p={P} q={Q} nlv={NLV} Float64 uniform weights scal=false "
                       "(BASELINE.json configs[1], README benchmark)"}
    if cpu:
        cfg["cpu_only"] = True
    else:
        cfg["parallelism"] = (f"rows sharded over {world} GPU(s), {n // world} rows each (strong scaling: n fixed); "
                              "one peer-HBM exchange of the packed Gram")
        cfg["l2"] = f"inputs ({8e-9 * (n // world) * (P + Q):.2f} GB per GPU) are larger than the 126 MB L2; no flush needed"
        cfg["vs_baseline_def"] = "value / (F_fit / 8.100469 s README plskern time on an i9-10885H)"
    return cfg


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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ our arm
def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_GLOBAL, help="global rows of the headline fit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + parity only (profiling runs)")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
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
    if world != args.gpus and rank == 0:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    gloo = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
        gloo = dist.new_group(backend="gloo")      # host-side barrier that leaves the GPUs idle
    dev.init(local_rank)
    dev.use_current_stream()
    lib = jc.lib()
    K, W = args.steps, max(args.warmup, 3)
    f64 = dict(dtype=torch.float64, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], **f64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, k):
        """k steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> ms per step"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        barrier()
        return allmax(e0.elapsed_time(e1)) / k

    # ---- the exchange: peer HBM windows (the product path) with NCCL timed beside it
    comm, comm_note = None, "single GPU: no exchange"
    if world > 1:
        try:
            comm = sharded.PeerComm(dev.packed_len(C4["p"], C4["q"]))
            comm_note = ("peer HBM windows over CUDA IPC (csrc/comm.cu), fused: K1b stores the reduced block into every "
                         "rank's slot and raises the flags, K3 reads the ordered sum of the slots")
        except Exception as ex:                     # IPC not permitted on this box: say so, use NCCL
            comm, comm_note = None, f"NCCL all-reduce (peer windows unavailable: {ex})"
        ok = torch.tensor([1.0 if comm is not None else 0.0], **f64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0.0 and comm is not None:
            comm.close()
            comm, comm_note = None, "NCCL all-reduce (peer windows unavailable on some rank)"
    use_comm = comm if args.exchange == "peer" else None
    if args.exchange == "nccl" and world > 1:
        comm_note = "NCCL all-reduce (--exchange nccl)"

    # ---- headline: strong scaling of C2 (n fixed), inputs generated on the device (bits of oracle/synth.py)
    n_glob = args.rows
    lo, hi = sharded.shard_rows(n_glob, rank, world)
    n_loc = hi - lo
    X, Y = dev.colmajor_empty_xy(max(n_loc, 2), P, Q, device)      # Y behind X: one n x (p + q) matrix for K1
    if n_loc > 0:
        dev.fill_uniform(X, n_loc, 1, lo, n_glob)
        dev.fill_uniform(Y, n_loc, 2, lo, n_glob)
    model = dev.DeviceModel(max(n_loc, 2), P, Q, NLV, device)
    pivot = torch.empty(P + Q + 1, **f64)
    packed = torch.empty(dev.packed_len(P, Q), **f64)

    def step():
        sharded.fit_sharded(X, Y, None, n_loc, model, scal=False, pivot=pivot, packed=packed, comm=use_comm)

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
    # the per-phase event records are instrumentation: out of the timed region (the event ring around K1, which the
    # roofline needs live from this region, stays on); the phase breakdown comes from one more step afterwards
    dev.set_phase_timing(False)
    t_wall0 = time.time()
    ms_step = timed(step, K)
    t_wall1 = time.time()
    launches = lib.jcb200_launch_count() - launches0
    gram_ms = _lib.gram_timings(K)
    dev.set_phase_timing(True)
    step()
    phases = dev.sync_timings()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = f_fit(n_glob, P, Q, NLV) / (ms_step * 1e-3) * 1e-12
    timeouts = lib.jcb200_comm_timeouts() if comm is not None else 0

    extras = not args.no_extras
    exchange = {"used": comm_note}
    if world > 1 and extras:
        # the same step with the other carrier, and the two exchanges alone on the C2 buffer
        def variant(cm, fused):
            def f():
                sharded.fit_sharded(X, Y, None, n_loc, model, scal=False, pivot=pivot, packed=packed, comm=cm,
                                    fused=fused)
            for _ in range(2):
                f()
            return timed(f, max(4, K // 2))
        exchange["step_ms"] = {"nccl_allreduce": variant(None, False)}
        if comm is not None:
            exchange["step_ms"]["peer_push_and_sum_kernels"] = variant(comm, False)
            exchange["step_ms"]["peer_fused_into_K1b_K3"] = variant(comm, True)
        buf = torch.zeros(dev.packed_len(P, Q), **f64)
        if comm is not None:
            for _ in range(3):
                comm.allreduce(buf)
            exchange["peer_allreduce_ms_c2"] = timed(lambda: comm.allreduce(buf), 20)
        for _ in range(3):
            dist.all_reduce(buf)
        exchange["nccl_allreduce_ms_c2"] = timed(lambda: dist.all_reduce(buf), 20)
        exchange["bytes_c2"] = buf.numel() * 8

    # ---- parity: the c2_cut golden (n = 20 000), fitted row-sharded over the N ranks, vs tests/golden/c2_cut.npz
    parity = golden_parity(np, torch, dist, dev, sharded, rank, world, device, use_comm)

    # ---- weak scaling (1e6 rows per GPU: round 1's headline), N > 1 only — at N = 1 it IS the headline
    weak = None
    if world > 1 and extras and n_glob == N_GLOBAL:
        del X, Y, model
        nw = N_GLOBAL
        Xw, Yw = dev.colmajor_empty_xy(nw, P, Q, device)
        dev.fill_uniform(Xw, nw, 1, rank * nw, nw * world)
        dev.fill_uniform(Yw, nw, 2, rank * nw, nw * world)
        mw = dev.DeviceModel(nw, P, Q, NLV, device)

        def step_w():
            sharded.fit_sharded(Xw, Yw, None, nw, mw, scal=False, pivot=pivot, packed=packed, comm=use_comm)
        for _ in range(2):
            step_w()
        ms_w = timed(step_w, max(4, K // 2))
        weak = {"scaling": "weak", "rows_per_gpu": nw, "n_global": nw * world, "ms_per_step": ms_w,
                "value": f_fit(nw * world, P, Q, NLV) / (ms_w * 1e-3) * 1e-12, "unit": "TFLOP/s"}
        del Xw, Yw, mw
        X = Y = model = None
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3]: n = 1e7, p = 2000, q = 10, nlv = 50 row-sharded over N >= 2 GPUs
    c4 = None
    if world > 1 and extras and not args.no_c4:
        c4 = c4_leg(torch, dist, dev, sharded, _lib, rank, world, device, use_comm, comm, timed, barrier)
        torch.cuda.empty_cache()

    # ---- end to end through the host-pointer C ABI
    if X is not None:
        del X, Y, model
        torch.cuda.empty_cache()
    Ke = max(1, min(args.e2e_steps, K))
    e2e = e2e_leg(np, torch, dist, jc, dev, sharded, rank, world, device, gloo, comm, Ke, extras)

    if rank != 0:
        if world > 1:
            dist.barrier(group=gloo)
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fp64 = peaks.get("fp64") or {}
    peak_tf = fp64.get("dmma_tflops_burst")
    gram_avg = sum(gram_ms) / len(gram_ms) if gram_ms else None
    achieved = f_gram(n_loc, P, Q) / (gram_avg * 1e-3) * 1e-12 if gram_avg else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic_r02.json")))["dram_bytes_per_launch"]
        traffic = traffic * n_loc / N_GLOBAL       # captured at 1e6 rows per launch
    except Exception:
        pass
    roofline = {
        "kernel": "gram_kernel<false> (K1: fused centre + DMMA SYRK/GEMM)", "bound": "tensor",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (achieved / peak_tf) if (achieved and peak_tf) else None, "traffic": traffic,
        "frac_of_cublas_dgemm_peak": (achieved / fp64["cublas_dgemm_8192_burst_tflops"])
        if (achieved and fp64.get("cublas_dgemm_8192_burst_tflops")) else None,
        "peak_source": f"measured FP64 DMMA.8x8x4 register-resident loop on this pool's B200 ({fp64.get('file')}; "
                       "MEASURED_PEAKS.json holds no FP64 figure)",
        "algorithmic_flops_per_launch": f_gram(n_loc, P, Q), "rows_per_launch": n_loc, "avg_launch_ms": gram_avg,
        "launches_timed": len(gram_ms), "share_of_step": (gram_avg / ms_step) if gram_avg else None,
    }
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and extras:      # the CPU leg is timed at N=1 only (rank 0)
        cpu_baseline, _ = cpu_baseline_record(3, 1, e2e.pop("_Xh", None), e2e.pop("_Yh", None))
    e2e.pop("_Xh", None)
    e2e.pop("_Yh", None)
    base_tf = f_fit(N_GLOBAL, P, Q, NLV) / README_PLSKERN_SECONDS * 1e-12
    line = {
        "metric": "plskern_fit_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_step, "fit_seconds": ms_step * 1e-3,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": (value / base_tf) if n_glob == N_GLOBAL else None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_glob, world),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": int(launches), "phases_ms_last_step": phases, "clocks": clocks,
        "exchange": exchange, "comm_timeouts": int(timeouts), "parity": parity, "weak": weak,
        "extra_configs": {"c4": c4} if c4 is not None else None,
    }
    emit(line)
    if world > 1:
        dist.barrier(group=gloo)
        dist.destroy_process_group()


def golden_parity(np, torch, dist, dev, sharded, rank, world, device, comm):
    """c2_cut (n=20000, p=500, q=10, nlv=25) fitted row-sharded over the ranks vs the committed golden:
    B for every k = 1..25 (from R, C and the scales), T on the golden's row sample, means; relative errors."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "c2_cut.npz"))
    except Exception as ex:
        return {"ok": False, "error": f"golden fixture unreadable: {ex}"}
    n, p, q, nlv = 20000, 500, 10, 25
    lo, hi = sharded.shard_rows(n, rank, world)
    nl = hi - lo
    Xg, Yg = dev.colmajor_empty_xy(max(nl, 2), p, q, device)
    if nl > 0:
        dev.fill_uniform(Xg, nl, 1, lo, n)
        dev.fill_uniform(Yg, nl, 2, lo, n)
    mg = dev.DeviceModel(max(nl, 2), p, q, nlv, device)
    sharded.fit_sharded(Xg, Yg, None, nl, mg, scal=False, comm=comm)
    torch.cuda.synchronize()
    R, C, W = mg.R[:nlv].T.cpu().numpy(), mg.C[:nlv].T.cpu().numpy(), mg.W[:nlv].T.cpu().numpy()   # p x a, q x a
    xs, ys = mg.xscales.cpu().numpy(), mg.yscales.cpu().numpy()
    s = np.sign(np.sum(z["W"] * W, axis=0))

    def rel(a, b):
        nb = np.linalg.norm(b)
        return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))
    errB = 0.0
    for k in range(1, nlv + 1):
        Bd = (R[:, :k] / xs[:, None]) @ C[:, :k].T * ys[None, :]
        Bg = (z["R"][:, :k] / z["xscales"][:, None]) @ z["C"][:, :k].T * z["yscales"][None, :]
        errB = max(errB, rel(Bd, Bg))
    errBk = max(rel((R[:, :k] / xs[:, None]) @ C[:, :k].T * ys[None, :], z["B_ks"][i])
                for i, k in enumerate(z["ks"]) if k > 0)
    # T on the golden's row sample: every rank grades the sample rows it holds
    rows = z["rows"]
    mine = (rows >= lo) & (rows < hi)
    num = den = 0.0
    if mine.any():
        Tl = mg.T[:nlv, :nl].cpu().numpy().T[rows[mine] - lo] * s
        num, den = float(np.sum((Tl - z["T_rows"][mine]) ** 2)), float(np.sum(z["T_rows"][mine] ** 2))
    t = torch.tensor([num, den], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t)
    errT = float((t[0] / t[1]).sqrt().item())
    out = {"vs": "tests/golden/c2_cut.npz (n=20000 p=500 q=10 nlv=25), rows sharded over the ranks",
           "B_every_k_1_25": errB, "B_golden_ks": errBk, "T_row_sample": errT,
           "R": rel(R * s, z["R"]), "C": rel(C * s, z["C"]), "TT": rel(mg.TT[:nlv].cpu().numpy(), z["TT"]),
           "xmeans": rel(mg.xmeans.cpu().numpy(), z["xmeans"]), "ymeans": rel(mg.ymeans.cpu().numpy(), z["ymeans"]),
           "tolerance": 1e-10}
    worst = max(v for k, v in out.items() if isinstance(v, float) and k != "tolerance")
    # every rank must hold the same model bits (replicated K3/K4 on a bit-identical reduced Gram)
    chk = torch.stack([mg.R[:nlv].sum(), mg.C[:nlv].sum(), mg.TT[:nlv].sum()])
    lo_, hi_ = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    out["ranks_bit_identical"] = bool(torch.equal(lo_, hi_))
    out["ok"] = bool(worst <= 1e-10 and out["ranks_bit_identical"])
    return out


def c4_leg(torch, dist, dev, sharded, _lib, rank, world, device, use_comm, comm, timed, barrier):
    """BASELINE configs[3]: n=1e7, p=2000, q=10, nlv=50, rows sharded over the N GPUs, inputs generated on the
    device by K8 (160 GB in total: never on the host).  Checked through size-independent properties."""
    n, p, q, nlv = C4["n"], C4["p"], C4["q"], C4["nlv"]
    lo, hi = sharded.shard_rows(n, rank, world)
    nl = hi - lo
    need = 8.0 * nl * (p + q + nlv + 2) + 3e9
    free, total = torch.cuda.mem_get_info(device)
    fits = torch.tensor([1.0 if need < free else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    if fits.item() == 0.0:
        return {"skipped": f"rank shard needs {need * 1e-9:.0f} GB, {free * 1e-9:.0f} GB free"}
    X, Y = dev.colmajor_empty_xy(nl, p, q, device)
    dev.fill_uniform(X, nl, 1, lo, n)
    dev.fill_uniform(Y, nl, 2, lo, n)
    model = dev.DeviceModel(nl, p, q, nlv, device)
    f64 = dict(dtype=torch.float64, device=device)
    pivot = torch.empty(p + q + 1, **f64)
    packed = torch.empty(dev.packed_len(p, q), **f64)

    def step():
        sharded.fit_sharded(X, Y, None, nl, model, scal=False, pivot=pivot, packed=packed, comm=use_comm)
    step()
    ksteps = 3
    ms = timed(step, ksteps)
    g = _lib.gram_timings(ksteps)
    ph = dev.sync_timings()
    gram_avg = sum(g) / len(g)
    out = {"workload": f"plskern fit n={n} p={p} q={q} nlv={nlv} Float64, {nl} rows per GPU on {world} GPUs "
                       "(BASELINE.json configs[3]), X generated on the device",
           "steps": ksteps, "ms_per_step": ms, "fit_seconds": ms * 1e-3,
           "tflops": f_fit(n, p, q, nlv) / (ms * 1e-3) * 1e-12,
           "k1_ms": gram_avg, "k1_tflops_per_gpu": f_gram(nl, p, q) / (gram_avg * 1e-3) * 1e-12,
           "phases_ms_last_step": ph, "hbm_gb_per_gpu": 8e-9 * nl * (p + q + nlv)}
    # the exchange alone on the 32 MB buffer, both carriers
    buf = torch.zeros(dev.packed_len(p, q), **f64)
    if comm is not None:
        for _ in range(2):
            comm.allreduce(buf)
        out["peer_allreduce_ms"] = timed(lambda: comm.allreduce(buf), 10)
    for _ in range(2):
        dist.all_reduce(buf)
    out["nccl_allreduce_ms"] = timed(lambda: dist.all_reduce(buf), 10)
    out["exchange_bytes"] = buf.numel() * 8
    # properties of a correct fit that need no oracle: ||w|| = 1, P'R = I, T'T / n = diag(TT) (uniform weights),
    # the model bits agree on every rank
    T = model.T[:nlv, :nl]
    G = T @ T.T
    dist.all_reduce(G)
    G = G / n
    TT = model.TT[:nlv]
    out["checks"] = {
        "w_norm_minus_1": float((model.W[:nlv].norm(dim=1) - 1).abs().max().item()),
        "PtR_minus_I": float((model.P[:nlv] @ model.R[:nlv].T - torch.eye(nlv, **f64)).abs().max().item()),
        "TtDT_minus_diagTT_rel": float(((G - torch.diag(TT)).abs().max() / TT.max()).item()),
    }
    chk = torch.stack([model.R[:nlv].sum(), model.C[:nlv].sum(), TT.sum()])
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    out["checks"]["ranks_bit_identical"] = bool(torch.equal(lo_, hi_))
    out["checks"]["ok"] = bool(out["checks"]["w_norm_minus_1"] < 1e-12 and out["checks"]["PtR_minus_I"] < 1e-9 and
                               out["checks"]["TtDT_minus_diagTT_rel"] < 1e-10 and out["checks"]["ranks_bit_identical"])
    del X, Y, model, T, G
    return out


def e2e_leg(np, torch, dist, jc, dev, sharded, rank, world, device, gloo, comm, Ke, extras):
    """The metric through the host-pointer C ABI, host<->device copies inside the timing (wall clock around the
    call a user makes).  N=1: one GPU.  N>1: rank 0 rebinds the library to all N GPUs (jcb200_init_multi) and
    makes the SAME call; the other ranks free their GPUs and wait on a host-side (gloo) barrier."""
    lib = jc.lib()
    if comm is not None:
        comm.close()
    if world > 1:
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        dist.barrier(group=gloo)
        if rank != 0:
            return {}
    n = N_GLOBAL
    h2d = 8 * (n * P + n * Q)
    d2h = 8 * (n * NLV + 3 * P * NLV + Q * NLV + NLV + 2 * P + 2 * Q + n)
    hX = torch.empty((P, n), dtype=torch.float64).pin_memory()
    hY = torch.empty((Q, n), dtype=torch.float64).pin_memory()
    # the bench inputs once more, from the device generator
    Xd = dev.colmajor_empty(n, P, device)
    Yd = dev.colmajor_empty(n, Q, device)
    dev.fill_uniform(Xd, n, 1, 0, n)
    dev.fill_uniform(Yd, n, 2, 0, n)
    hX.copy_(Xd[:, :n])
    hY.copy_(Yd[:, :n])
    torch.cuda.synchronize()
    del Xd, Yd
    torch.cuda.empty_cache()
    Xh, Yh = hX.numpy().T, hY.numpy().T          # column-major [n, p] views of the pinned buffers
    assert Xh.flags.f_contiguous and Yh.flags.f_contiguous
    ndev = 1
    if world > 1:
        if torch.cuda.device_count() >= world:
            lib.jcb200_shutdown()
            jc.init_multi(list(range(world)))
            ndev = world
        else:
            ndev = 1                              # the ranks cannot see each other's GPUs: one-GPU call
    else:
        dev.use_own_stream()                      # the C ABI call times itself on its own stream

    def wall(fn, k, warm):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        return (time.perf_counter() - t0) / k

    fit = lambda: jc.plskern(Xh, Yh, nlv=NLV)
    sec = wall(fit, Ke, 3)
    phases = jc.last_timings()
    out = {"value": f_fit(n, P, Q, NLV) / sec * 1e-12, "unit": "TFLOP/s", "fit_seconds": sec,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke, "devices": ndev,
           "how": ("jcb200_plskern_fit (C ABI, host pointers) on page-locked host arrays, wall clock"
                   + (f"; library bound to {ndev} GPUs (jcb200_init_multi): rows sharded inside the call, one PCIe "
                      "link per GPU, peer-memory Gram reduce" if ndev > 1 else "")),
           "phases_ms": phases, "vs_readme_seconds": README_PLSKERN_SECONDS / sec}
    if extras:
        # what an ordinary (pageable) array costs — a Julia `rand(n, p)` handed over as is; T still comes from
        # the library's page-locked pool (as JchemoB200.jl and the Python mirror allocate it)
        Xp, Yp = np.array(Xh, order="F"), np.array(Yh, order="F")
        secp = wall(lambda: jc.plskern(Xp, Yp, nlv=NLV), max(2, Ke // 2), 1)
        out["pageable"] = {"fit_seconds": secp, "value": f_fit(n, P, Q, NLV) / secp * 1e-12,
                           "how": "same call, X and Y in ordinary pageable memory (threaded staging through two "
                                  "page-locked 32 MB slots)"}
        # plskern! (README's second number, 7.23 s): the centred X and Y travel back (4.08 GB more D2H)
        Xb, Yb = hX.numpy().T, hY.numpy().T

        def bang():
            jc.plskern_bang(Xb, Yb, nlv=NLV)      # repeated in place: the data stay centred, the work is the same
        secb = wall(bang, max(2, Ke // 2), 1)
        out["plskern_bang"] = {"fit_seconds": secb, "value": f_fit(n, P, Q, NLV) / secb * 1e-12,
                               "d2h_bytes_per_step": d2h + h2d, "phases_ms": jc.last_timings(),
                               "vs_readme_seconds": README_PLSKERN_BANG_SECONDS / secb,
                               "how": "jcb200_plskern_fit(writeback_xy=1) on page-locked arrays: K5 + K7 in 8 row "
                                      "blocks, each block's T / X / Y copies under the next blocks' kernels"}
        hX.copy_(torch.from_numpy(Xp.T))          # restore the raw inputs
        hY.copy_(torch.from_numpy(Yp.T))
        if ndev == 1:
            # device-resident data handle: fit + summary + gridscore on ONE upload of X (second pass timed: the
            # first one pays one-off costs — pandas import, workspace growth, K1 schedules of the new shapes)
            def workflow():
                t0 = time.perf_counter()
                fm = jc.plskern(Xh, Yh, nlv=NLV)
                t1 = time.perf_counter()
                jc.summary(fm, Xh)
                t2 = time.perf_counter()
                jc.gridscorelv(Xh, Yh, Xh, Yh, score="rmsep", nlv=range(0, NLV + 1))
                t3 = time.perf_counter()
                return t1 - t0, t2 - t1, t3 - t2
            workflow()
            plain = workflow()                     # every call uploads X again
            t0 = time.perf_counter()
            with jc.resident(Xh, Yh):
                up = time.perf_counter() - t0
                workflow()
                res = workflow()
            out["resident"] = {"upload_seconds": up, "fit_seconds": res[0], "summary_seconds": res[1],
                               "gridscorelv_seconds": res[2],
                               "same_calls_without_handle_seconds": {"fit": plain[0], "summary": plain[1],
                                                                     "gridscorelv": plain[2]},
                               "how": "jcb200_resident_add(X), (Y) once; plskern, summary, gridscorelv (a fit + the "
                                      "26-nlv scoring sweep) then run without any transfer of X"}
        del Xp, Yp
    out["_Xh"], out["_Yh"] = Xh, Yh
    return out


if __name__ == "__main__":
    main()
