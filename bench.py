#!/usr/bin/env python
"""Benchmark of the batched SRBD-MPC solve (BASELINE.json metric):

    MPC solves/sec at 4096 envs (form + 5-iteration PDIPM, cold start) per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: oracle restatement

A "step" is one pass of the hot path over one batch of synthetic randomized biped states and
gait schedules (biped-pympc_b200/workload.py): `qp_former` assembly +
`sparse_pdipm_multiple_iterations` (5 Newton iterations from the reference cold start) for
every env of the batch, through the fused C-ABI entry `srbd_mpc_solve`.
Weak scaling: every rank owns its own 4096-env batch (envs are independent; no collective on
the solve path), value = envs all ranks solved / max-over-ranks device time.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
ITERS = 5
FLOP_PER_ITER = 170e3          # structured count per Newton iteration per env
BYTES_ASSEMBLY = 21664         # 3 616 B in + 18 048 B out per env
METRIC = "mpc_solves_per_sec_4096_envs_form_plus_5iter_pdipm"
UNIT = "solves/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms from before the warm-up; stop(t0, t1)
    keeps the samples whose timestamp falls inside the timed region (all of them if the region was
    shorter than one sampling period, which is then said in the result)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, l in rows:
                p = [x.strip() for x in l.split(",")]
                if len(p) < 10:
                    continue
                try:
                    sm.append(float(p[2])); mx.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        inside = [r for r in self.lines if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not inside:
            inside, window = self.lines, "whole run (timed region shorter than one 20 ms sampling period)"
        sm, mx, reasons = parse(inside)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference_run(n_envs, iters, steps, warmup, threads, seed=20261018):
    """Times the reference's CPU path on a bounded sample of the workload.  Preferred: oracle/_ref/lib<fn>_cpu.so, the
    expression tapes of the reference's two functions (built from its sources by oracle/make_cusadi_ref.py) compiled to C
    and run with `omp parallel for` over envs, the reference's own CPU convention (evaluate_parallel_cpu.cpp:79-82) --
    kind "reference".  Otherwise the C restatement oracle/srbd_oracle.c (literal qp_former + full-KKT sparse LDL^T
    PDIPM, OpenMP over envs) -- kind "port".  Returns (solves/s, per-step seconds, kind, description)."""
    import biped_pympc_b200 as bp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cusadi_ref
    inp, _ = bp.workload.make_batch(n_envs, seed=seed)
    if cusadi_ref.cpu_available() and iters == 5:
        cpu = cusadi_ref.CusadiRefCpu()

        def one():
            cpu.pdipm5(cpu.qp_former(inp, threads), threads=threads)
        kind = "reference"
        desc = ("oracle/_ref: expression tapes of the reference's qp_former and sparse_pdipm_multiple_iterations (its sources run over a "
                "CasADi stand-in; CasADi itself is not installable offline) compiled to C with gcc -O1, omp parallel for over envs")
    else:
        from oracle import c_oracle as co

        def one():
            co.pdipm(co.qp_former(inp, nthreads=threads), iters, nthreads=threads)
        kind = "port"
        desc = "oracle/srbd_oracle.c (port of qp_former + full-KKT sparse LDL^T PDIPM, OpenMP over envs)"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_envs / (sum(times) / len(times)), times, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 2048
    value, times, kind, desc = cpu_reference_run(sample, ITERS, args.steps, max(args.warmup, 1), threads)
    ms = 1e3 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{ENVS_PER_GPU}-env batched SRBD MPC, 5 PDIPM iters, fp64 (BASELINE configs[1])",
                   "envs_per_step_sample": sample, "iters": ITERS, "horizon": 10},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} envs/step of the same seeded workload; {desc}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_cusadi_ref(args, emit=True):
    """`--impl cusadi`: the reference's CusADi GPU path on this GPU -- the two kernels the reference's own
    generateCUDACode.py emits (oracle/_ref, built by oracle/make_cusadi_ref.py), called the way CusadiFunction /
    MPCControllerCusadi call them: outputs and the (B, sz_w) work tensor are zero-filled before each launch
    (CusadiFunction.py:104-109), one thread per env runs the whole expression tape, synchronous launches.  A step =
    qp_former + the cold start + ONE sparse_pdipm_multiple_iterations call (5 iterations), like the default arm.
    Not part of the driver contract; it answers north_star's "next to the reference's CusADi GPU path"."""
    import torch
    import biped_pympc_b200 as bp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cusadi_ref
    var = cusadi_ref.solver_variant()
    if var is None or not cusadi_ref.available("qp_former"):
        line = {"impl": "cusadi", "unavailable": "oracle/_ref CusADi kernels not built (python oracle/make_cusadi_ref.py)"}
        if emit:
            print(json.dumps(line))
        return line
    suffix, per_call = var
    dev, B = "cuda:0", args.envs
    inp, _ = bp.workload.make_batch(B, seed=20261018)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    former = cusadi_ref.CusadiRefFunction("qp_former", B, dev)
    solver = cusadi_ref.CusadiRefFunction("sparse_pdipm_multiple_iterations", B, dev, suffix)
    kern = {"qp_former": [], "solver": []}

    def step():
        kern["qp_former"].append(former.evaluate(ins))
        Hv, f, Av, b, Gv, d = former.outputs_sparse
        x = torch.zeros(B, 240, dtype=torch.float64, device=dev)          # mpc_controller_cusadi.py:138-141
        s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=dev)
        t = 0.0
        for _ in range(ITERS // per_call):      # one call of the 5-iteration kernel, or the one-iteration build chained
            t += solver.evaluate([Hv, Gv, Av, f, d, b, x, s, z, y])
            if per_call != ITERS:
                x, s, z, y = [o.clone() for o in solver.outputs_sparse[:4]]
        kern["solver"].append(t)

    for _ in range(max(args.warmup, 1)):
        step()
    kern = {"qp_former": [], "solver": []}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    mean = lambda v: 1e3 * sum(v) / len(v)  # noqa: E731
    line = ({
        "impl": "cusadi", "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{B}-env batched SRBD MPC, 5 PDIPM iters, fp64 (BASELINE configs[1]) through the reference's CusADi-generated "
                               "kernels (its generateCUDACode.py over its function sources; expression graph from the CasADi stand-in)",
                   "envs_per_gpu": B, "iters": ITERS,
                   "solver_kernel": ("the reference's 5-iteration function in one kernel" if per_call == ITERS else
                                     "the reference's function built with num_iters=1, chained 5 x (ptxas needs > 58 GB for the 5-iteration kernel)")},
        "kernel_ms": {"qp_former": mean(kern["qp_former"]), "sparse_pdipm_multiple_iterations": mean(kern["solver"])},
        "instructions_per_env": {"qp_former": former.n_instructions, "sparse_pdipm_multiple_iterations": solver.n_instructions},
        "work_doubles_per_env": {"qp_former": former.sz_w, "sparse_pdipm_multiple_iterations": solver.sz_w}})
    if emit:
        print(json.dumps(line))
    return line


def run_b200(args):
    import numpy as np
    import torch
    import biped_pympc_b200 as bp

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the
        # ONE JSON line by pointing fd 1 at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.envs
    inp, rot = bp.workload.make_batch(B, seed=20261018, rank=rank)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
    ctx = bp.Context(B, local_rank)
    # L2 flush buffer (> 126 MB L2) written between timed steps
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB

    iters_done = torch.zeros(B, dtype=torch.int32, device=dev) if args.tol > 0 else None

    def step():
        return ctx.solve(ins, rot_t, num_calls=args.calls, iters_per_call=args.iters, tol=args.tol, iters_done=iters_done)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    for a, b in ev:
        flush.fill_(1.0)           # evict L2 between timed iterations (not timed)
        a.record()
        step()
        b.record()
    barrier()
    t_wall1 = time.time()
    times_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(times_ms)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI (H2D + kernels + D2H inside the timed region) ----
    # x, u (inputs 1, 2) are the linearisation point, all ones at the reference call site
    # (mpc_controller_cusadi.py:55-56) and dead for this LTV-linear problem: the host API takes
    # None for them and does not copy them
    assert all(float(inp[i].min()) == 1.0 == float(inp[i].max()) for i in (1, 2))
    h_in = [None if i in (1, 2) else torch.as_tensor(a).pin_memory() for i, a in enumerate(inp)]
    h_rot = torch.as_tensor(rot, dtype=torch.float32).pin_memory()
    h_w = torch.empty((B, 2, 6), dtype=torch.float32).pin_memory()
    for _ in range(3):
        ctx.solve_host(h_in, h_rot, h_w, num_calls=args.calls, iters_per_call=args.iters, tol=args.tol)
    n_e2e = max(10, min(args.steps, 100))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        ctx.solve_host(h_in, h_rot, h_w, num_calls=args.calls, iters_per_call=args.iters, tol=args.tol)   # synchronous
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1)
    e2e_value = world * B * n_e2e / e2e_s
    h2d = sum(a.numel() * 8 for a in h_in if a is not None) + h_rot.numel() * 4
    d2h = h_w.numel() * 4

    default_wl = (args.calls == 1 and args.iters == ITERS and args.tol <= 0)
    line = {
        "metric": METRIC if default_wl else "mpc_solves_per_sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"{B}-env batched SRBD MPC per GPU, 5 PDIPM iters, fp64, cold start "
                                "(BASELINE configs[1]); form + solve + wrench fused in one kernel via srbd_mpc_solve") if default_wl else
                               (f"{B}-env batched SRBD MPC per GPU, {args.calls} x {args.iters} PDIPM iters"
                                + (f" or tol {args.tol:g}" if args.tol > 0 else "") + ", fp64, cold start; srbd_mpc_solve"),
                   "envs_per_gpu": B, "iters": args.calls * args.iters, "horizon": 10, "parallelism": f"env-sharded x{world}, no collective",
                   "l2": "256 MB buffer written between timed iterations (L2 flush)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "api": "srbd_mpc_solve_host (pinned host buffers read and written over PCIe by the kernels themselves: the first 2058-env chunk in place, the second pulled into device staging by a co-resident prefetch kernel while the first solves)"},
        "gpu_launches": args.steps,          # ONE kernel per step: fused assembly + PDIPM + wrench (pdipm_fast_kernel)
        "clocks": clocks,
    }
    ts = sorted(times_ms)
    line["latency_ms"] = {"p50": ts[len(ts) // 2], "p99": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "max": ts[-1]}

    if iters_done is not None:
        hist = torch.bincount(iters_done.clamp(min=0).to(torch.int64), minlength=args.calls * args.iters + 1).tolist()
        line["iterations_histogram"] = {str(i): c for i, c in enumerate(hist) if c}
    n_it = args.calls * args.iters
    if rank == 0:
        # ---- roofline of the dominant kernel (PDIPM), timed alone with CUDA events -------------
        peaks, peak_src = _peaks()
        qp = bp.qp_form(ins)
        Hv, f, Av, bvec, Gv, d = qp
        outs = [torch.empty((B, n), dtype=torch.float64, device=dev) for n in bp._native.IP_NNZ_OUT]
        work = torch.empty((B, bp._native.WORK_DOUBLES), dtype=torch.float64, device=dev)
        reps = 20
        k_ms = []
        for i in range(reps + 3):
            flush.fill_(0.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            bp.pdipm([Hv, Gv, Av, f, d, bvec], num_iters=n_it, tol=args.tol, cold_start=True, stage_invariant=True, out=outs, work=work,
                     iters_done=iters_done)
            b.record()
            b.synchronize()
            if i >= 3:
                k_ms.append(a.elapsed_time(b))
        pd_ms = sum(k_ms) / len(k_ms)
        f_ms = []
        for i in range(reps + 3):
            flush.fill_(0.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            bp.qp_form(ins, out=qp)
            b.record()
            b.synchronize()
            if i >= 3:
                f_ms.append(a.elapsed_time(b))
        fm_ms = sum(f_ms) / len(f_ms)
        dfma = max(bp.dfma_peak(), bp.dfma_peak())
        done_iters = float(iters_done.clamp(min=0).sum().item()) if iters_done is not None else float(B * n_it)
        ach = done_iters * FLOP_PER_ITER / (pd_ms * 1e-3) / 1e12
        traffic, cap = None, {}
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_capture.json")) as fcap:
                cap = json.load(fcap)
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        except Exception:
            pass
        line["roofline"] = {"kernel": "pdipm_fast_kernel<14> (stage-invariant PDIPM, one warp per env)", "bound": "fp64",
                            "achieved": ach, "peak": dfma,
                            "unit": "TFLOP/s", "frac": ach / dfma, "traffic": traffic,
                            "traffic_note": "DRAM bytes per launch from the committed ncu capture "
                                            f"({cap.get('capture')}); algorithmic I/O of the fused tick is 1 696 B of inputs + 5 640 B of iterates + 48 B of wrench per env = {7384 * B} B: "
                                            "the kernel is bound by the FP64 pipe and shared-memory bandwidth, not HBM",
                            "smem_wavefronts_per_sm_cycle_ncu": cap.get("smem_wavefronts_per_sm_cycle"),
                            "kernel_ms": pd_ms, "share_of_step": pd_ms / (total_ms / args.steps) if world == 1 else None,
                            "peak_source": "DFMA micro-benchmark measured in this run (srbd_dfma_peak); "
                                           "MEASURED_PEAKS.json has no FP64 entry",
                            "algorithmic": f"{FLOP_PER_ITER:.0f} flop/iter/env x {done_iters / B:.2f} iters (mean) x {B} envs"}
        # The resource that actually bounds the kernel (north_star: "FP64-pipe or smem-bandwidth utilisation for PDIPM"): the
        # shared-memory data pipe moves one 128-byte wavefront per SM per clock.  Wavefronts per env-iteration come from the
        # committed ncu capture (they do not depend on the clock or the batch), the time and the clock are this run's.
        wf = cap.get("smem_wavefronts_per_launch")
        clk = clocks or {}
        if wf and clk.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(0).multi_processor_count
            per_env_iter = wf / (4096 * 5)
            ach_wf = per_env_iter * done_iters / (pd_ms * 1e-3) / 1e9
            peak_wf = sms * clk["sm_mhz"] * 1e6 / 1e9
            line["roofline_smem"] = {"kernel": "pdipm_fast_kernel<14>", "bound": "smem", "achieved": ach_wf, "peak": peak_wf,
                                     "unit": "Gwavefront/s (128 B each)", "frac": ach_wf / peak_wf,
                                     "wavefronts_per_env_iteration": per_env_iter,
                                     "peak_source": f"{sms} SMs x 1 wavefront/clk x the SM clock sampled in this run ({clk['sm_mhz']:.0f} MHz)",
                                     "source": cap.get("capture")}
        ach_bw = B * BYTES_ASSEMBLY / (fm_ms * 1e-3) / 1e9
        line["roofline_assembly"] = {"kernel": "qp_form_kernel", "bound": "hbm", "achieved": ach_bw,
                                     "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                     "frac": ach_bw / peaks.get("hbm_gbs", 6650.0),
                                     "traffic": cap.get("assembly_dram_bytes_per_launch"),
                                     "traffic_note": "DRAM bytes of one launch from profiles/ncu_qpform_r1_4096_summary.csv: below the "
                                                     f"{B * BYTES_ASSEMBLY} algorithmic bytes because most of the output is still in L2 at kernel end",
                                     "kernel_ms": fm_ms, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"}
        # ---- CPU baseline on the box's host cores (bounded sample) -----------------------------
        if world == 1 and default_wl:
            threads = os.cpu_count() or 1
            sample = 2048
            cpu_val, _, kind, desc = cpu_reference_run(sample, ITERS, 3, 1, threads)
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{sample} envs x 3 steps of the same workload; {desc}"}
            # the reference's CusADi GPU path on this GPU, same workload (north_star: "next to the reference's CusADi GPU
            # path where it can be compiled offline"); see run_cusadi_ref / oracle/make_cusadi_ref.py for what it is
            try:
                import argparse as _ap
                cu = run_cusadi_ref(_ap.Namespace(envs=B, steps=3, warmup=2), emit=False)
                if "unavailable" in cu:
                    line["cusadi_gpu_baseline"] = {"unavailable": cu["unavailable"]}
                else:
                    line["cusadi_gpu_baseline"] = {"value": cu["value"], "unit": UNIT, "ms_per_step": cu["ms_per_step"],
                                                   "kernel_ms": cu["kernel_ms"], "speedup_of_this_path": value / cu["value"],
                                                   "what": cu["config"]["workload"] + "; " + cu["config"]["solver_kernel"]}
            except Exception as exc:      # a missing / broken baseline library must not take the bench line down
                line["cusadi_gpu_baseline"] = {"unavailable": repr(exc)}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if saved_stdout is not None:
            os.dup2(2, 1)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_closed_loop(args):
    """Config 5 (not the default bench line): one JSON line with per-step latency percentiles of the stack."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import closed_loop_stack
    res = closed_loop_stack.main(B=args.envs, steps=args.steps, decimation=args.decimation, mpc_calls=args.calls if args.calls > 1 else 4,
                                 iters_per_call=args.iters, warmup=args.warmup, graphs=args.graphs)
    print(json.dumps({"metric": "closed_loop_env_control_steps_per_sec", "value": res["env_control_steps_per_s"], "unit": "env-steps/s",
                      "n_gpus": 1, "higher_is_better": True, "dtype": "f64 solve / f32 stack", "data": "synthetic",
                      "config": {"workload": "closed-loop rollout: gait generator + state estimator + HECTOR leg / swing-leg controllers + MPC "
                                             "(BASELINE configs[4]); SRBD plant with kinematic legs as stand-in simulator, controller calls timed",
                                 **{k: res[k] for k in ("envs", "control_steps", "decimation", "mpc_iterations_per_tick", "cuda_graphs")}},
                      **{k: v for k, v in res.items() if k not in ("envs", "control_steps", "decimation", "mpc_iterations_per_tick", "cuda_graphs", "env_control_steps_per_s")}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cusadi"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    # other BASELINE configs (not the default bench line): --calls 4 = the reference tick (4 x 5 iterations,
    # mpc_controller_cusadi.py:144-169); --envs 16384 --iters 30 --tol 1e-8 = run to convergence (config 4)
    ap.add_argument("--calls", type=int, default=1, help="chained PDIPM calls per tick")
    ap.add_argument("--iters", type=int, default=ITERS, help="Newton iterations per call")
    ap.add_argument("--tol", type=float, default=0.0, help="> 0: stop an env when rx, rs, re, mu < tol")
    # BASELINE configs[4]: closed-loop rollout of the whole control stack (tests/closed_loop_stack.py); --steps =
    # control steps at 1 kHz, --decimation = control steps per MPC tick (10 = the reference's, 1 = MPC every step)
    ap.add_argument("--closed-loop", action="store_true")
    ap.add_argument("--decimation", type=int, default=10)
    ap.add_argument("--graphs", type=int, default=1, help="closed loop: replay the control step as a CUDA graph")
    args = ap.parse_args()
    if args.closed_loop:
        run_closed_loop(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "cusadi":
        if int(os.environ.get("RANK", "0")) == 0:
            run_cusadi_ref(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
