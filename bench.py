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


def cpu_reference_run(n_envs, iters, steps, warmup, threads, seed=20261018, want="reference"):
    """Times the reference's CPU path on the workload.  want="reference": oracle/_ref/lib<fn>_cpu*.so, the expression tapes
    of the reference's two functions (built from its sources by oracle/make_cusadi_ref.py) compiled to C and run with
    `omp parallel for` over envs, the reference's own CPU convention (evaluate_parallel_cpu.cpp:79-82; gcc -O3
    -march=x86-64-v3 -- run_benchmarks.py:46 uses -O3 -march=native -- where that build exists, else the -O1 build) -- kind "reference".
    want="port" (or no oracle/_ref): the C restatement oracle/srbd_oracle.c (literal qp_former + full-KKT sparse LDL^T
    PDIPM, gcc -O3 -march=native, OpenMP over envs) -- kind "port".
    Returns (solves/s, per-step seconds, kind, description)."""
    import biped_pympc_b200 as bp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cusadi_ref
    inp, _ = bp.workload.make_batch(n_envs, seed=seed)
    if want == "reference" and cusadi_ref.cpu_available() and iters == 5:
        cpu = cusadi_ref.CusadiRefCpu()

        def one():
            cpu.pdipm5(cpu.qp_former(inp, threads), threads=threads)
        kind = "reference"
        desc = ("oracle/_ref: expression tapes of the reference's qp_former and sparse_pdipm_multiple_iterations (its sources run over a "
                "CasADi stand-in; CasADi itself is not installable offline) compiled to C, omp parallel for over envs; compiler flags: "
                + cpu.flags)
    else:
        from oracle import c_oracle as co

        def one():
            co.pdipm(co.qp_former(inp, nthreads=threads), iters, nthreads=threads)
        kind = "port"
        desc = "oracle/srbd_oracle.c (port of qp_former + full-KKT sparse LDL^T PDIPM, gcc -O3 -march=native, OpenMP over envs)"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_envs / (sum(times) / len(times)), times, kind, desc


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores, on the metric's own
    batch (4 096 envs per step, the same seeded workload as the GPU arm), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.envs
    # one step is ~0.1-0.2 s of host time: K steps as asked, capped so that the run ends within a few minutes
    steps = max(1, min(args.steps, 600))
    value, times, kind, desc = cpu_reference_run(B, ITERS, steps, max(min(args.warmup, 5), 1), threads)
    port_value, _, _, port_desc = cpu_reference_run(B, ITERS, min(steps, 5), 1, threads, want="port")
    ms = 1e3 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"{B}-env batched SRBD MPC per GPU, 5 PDIPM iters, fp64, cold start "
                                "(BASELINE configs[1]); form + solve + wrench fused in one kernel via srbd_mpc_solve"),
                   "envs_per_gpu": B, "iters": ITERS, "horizon": 10, "parallelism": "host threads, omp parallel for over envs",
                   "steps_timed": steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"the metric's full batch: {B} envs/step of the same seeded workload, {steps} steps; {desc}"},
        "cpu_baseline_port": {"value": port_value, "unit": UNIT, "cores": threads, "kind": "port",
                              "sample": f"{B} envs/step, {min(steps, 5)} steps; {port_desc}"},
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


# ----------------------------------------------------------------------------------------------------------------------
# Sub-records of the default line: the other BASELINE.json configs, each bounded to a few seconds
# ----------------------------------------------------------------------------------------------------------------------
def sub_strong_scaling(bp, torch, dist, dev, local_rank, rank, world, flush, steps=10):
    """BASELINE configs[2] at fixed total size: 65 536 envs over the N GPUs of the run (65 536 / N per rank), one fused
    launch per rank, no collective on the solve path; time = max over ranks.  Run by every rank."""
    total = 65536
    Bs = total // world
    inp, rot = bp.workload.make_batch(Bs, seed=20261018 + 1000, rank=rank)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
    ctx = bp.Context(Bs, local_rank)
    for _ in range(3):
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1.0)
        a.record()
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.close()
    del ins, rot_t
    return {"value": total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "envs_total": total, "envs_per_gpu": Bs, "n_gpus": world,
            "steps": steps, "scaling": "strong", "what": "BASELINE configs[2]: 65 536 envs sharded over the GPUs of this run, 5 PDIPM iterations"}


def sub_converge(bp, torch, dev, local_rank, flush, steps=3):
    """BASELINE configs[3]: 16 384 envs to tol 1e-8 on (|rx|, |rs|, |re|, mu), at most 30 iterations."""
    B = 16384
    inp, rot = bp.workload.make_batch(B, seed=20261018 + 2000)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
    itd = torch.zeros(B, dtype=torch.int32, device=dev)
    ctx = bp.Context(B, local_rank)
    ctx.solve(ins, rot_t, num_calls=1, iters_per_call=30, tol=1e-8, iters_done=itd)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1.0)
        a.record()
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=30, tol=1e-8, iters_done=itd)
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    hist = torch.bincount(itd.clamp(min=0).to(torch.int64), minlength=31).tolist()
    # a looser tolerance for comparison: the iteration counts then differ between envs
    ctx.solve(ins, rot_t, num_calls=1, iters_per_call=30, tol=1e-2, iters_done=itd)
    torch.cuda.synchronize()
    hist2 = torch.bincount(itd.clamp(min=0).to(torch.int64), minlength=31).tolist()
    ctx.close()
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "envs": B, "max_iters": 30, "tol": 1e-8, "steps": steps,
            "iterations_histogram": {str(i): c for i, c in enumerate(hist) if c},
            "iterations_histogram_tol_1e-2": {str(i): c for i, c in enumerate(hist2) if c},
            "what": "BASELINE configs[3]: run to convergence, beyond the reference's 5-iteration codegen limit"}


def sub_single_env(bp, torch, dev, local_rank, reps=200):
    """BASELINE configs[0]: the single T1 standing problem, batch = 1 (the reference's CPU-runnable case): latency of one
    solve on the GPU (one warp of one SM) next to the reference's own tape on one host core."""
    inp = bp.workload.config1_t1_standing()
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.eye(3, dtype=torch.float32, device=dev).reshape(1, 3, 3).contiguous()
    ctx = bp.Context(1, local_rank)
    for _ in range(20):
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    t0 = time.perf_counter()
    for _ in range(50):
        ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50
    ctx.close()
    out = {"envs": 1, "iters": ITERS, "gpu_ms": {"p50": ts[len(ts) // 2], "p99": ts[int(0.99 * (len(ts) - 1))]},
           "gpu_call_plus_sync_wall_ms": 1e3 * wall,
           "what": "BASELINE configs[0]: single T1 biped standing problem, horizon 10, 5 PDIPM iterations, batch = 1 (CUDA events "
                   "around one call / wall clock of call + synchronize)"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cusadi_ref
        if cusadi_ref.cpu_available():
            cpu = cusadi_ref.CusadiRefCpu()
            for _ in range(3):
                cpu.pdipm5(cpu.qp_former(inp, 1), threads=1)
            t0 = time.perf_counter()
            for _ in range(20):
                cpu.pdipm5(cpu.qp_former(inp, 1), threads=1)
            out["cpu_reference_tape_ms"] = 1e3 * (time.perf_counter() - t0) / 20
            out["cpu_flags"] = cpu.flags
    except Exception as exc:
        out["cpu_reference_tape_ms"] = None
        out["cpu_error"] = repr(exc)
    return out


def sub_fp32_mode(bp, torch, dev, local_rank, flush, ref_ms, steps=10):
    """north_star's optional fp32 mode (include/srbd_mpc.h SRBD_F32_FACTOR / _REFINE / _REFINE2), measured on the metric's
    4 096-env workload: time of the fused tick per mode and the worst per-env relative error of x after the 5 iterations
    against this repo's own fp64 path (which the parity tests hold within 1e-9 of the 80-bit oracle)."""
    B = 4096
    inp, rot = bp.workload.make_batch(B, seed=20261018)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
    ctx = bp.Context(B, local_rank)
    x64 = ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS, want_iterates=True)[1]
    out = {"what": "optional reduced-precision mode: fp32 Schur pivot blocks (formed, coupled, inverted, stored in fp32) with 0 / 1 / 2 "
                   "steps of iterative refinement per solve against the fp64 operator; err_x = max over envs of |x - x_fp64| / |x_fp64| "
                   "after 5 iterations; stated tolerance 1e-4 = the two-step mode",
           "fp64_ms_per_step": ref_ms, "modes": {}}
    for mode in ("f32", "f32_refined", "f32_refined2"):
        ctx.set_precision(mode)
        x = ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS, want_iterates=True)[1]
        err = float(((x - x64).norm(dim=1) / x64.norm(dim=1)).max().item())
        for _ in range(2):
            ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.fill_(1.0)
            a.record()
            ctx.solve(ins, rot_t, num_calls=1, iters_per_call=ITERS)
            b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        out["modes"][mode] = {"ms_per_step": ms, "value": B / (ms * 1e-3), "unit": UNIT, "err_x": err, "speed_vs_fp64": ref_ms / ms}
    ctx.close()
    return out


def sub_closed_loop():
    """BASELINE configs[4]: gait generator + swing leg + MPC at 4 096 envs for 1 000 control steps (tests/closed_loop_stack.py:
    BipedControllerB200 in closed loop with a single-rigid-body stand-in plant; only the controller calls are timed)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import closed_loop_stack
    res = closed_loop_stack.main(B=4096, steps=1000, decimation=10, mpc_calls=4, iters_per_call=5, warmup=20, graphs=1)
    res["what"] = ("BASELINE configs[4]: 4 096 envs x 1 000 control steps at 1 kHz, MPC tick (4 x 5 iterations) every 10th step, "
                   "control step replayed as a CUDA graph; latencies in ms per control step")
    return res


class _CompatFunction:
    """CusadiFunction.py:28-109 restated for measurement: the calling convention an UNMODIFIED CusadiFunction applies to
    lib<fn>.so::evaluate -- zero-filled outputs and (B, sz_w) work tensor before every call (:104-109), the input
    pointer table rewritten element by element (:97-102), device arrays of device pointers, and getDenseOutput through
    torch.sparse_coo_tensor (:49-58).  `libdir` selects whose shared objects are loaded."""

    def __init__(self, spec, B, libdir, sz_w, dev, suffix=""):
        import ctypes
        import torch
        self.fn, self.B, self.dev, self.torch = spec, B, dev, torch
        self.lib = ctypes.CDLL(os.path.join(libdir, f"lib{spec.name()}{suffix}.so"))
        self.lib.evaluate.restype = ctypes.c_float
        self.lib.evaluate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        f64 = dict(dtype=torch.float64, device=dev)
        self.outputs_sparse = [torch.zeros(B, spec.nnz_out(i), **f64) for i in range(spec.n_out())]
        self.work = torch.zeros(B, sz_w, **f64)
        self.in_ptrs = torch.zeros(spec.n_in(), dtype=torch.int64, device=dev)
        self.out_ptrs = torch.zeros(spec.n_out(), dtype=torch.int64, device=dev)
        for i, t in enumerate(self.outputs_sparse):
            self.out_ptrs[i] = t.data_ptr()
        self.kernel_s = 0.0
        self._trip = {i: [self.torch.as_tensor(v) for v in spec.triplet_out(i)] for i in range(spec.n_out())}

    def evaluate(self, inputs):
        for o in self.outputs_sparse:
            o.zero_()
        self.work.zero_()
        self._keep = [t.contiguous() for t in inputs]
        for i, t in enumerate(self._keep):
            self.in_ptrs[i] = t.data_ptr()
        self.kernel_s += float(self.lib.evaluate(self.in_ptrs.data_ptr(), self.work.data_ptr(), self.out_ptrs.data_ptr(), self.B))

    def getDenseOutput(self, i):
        torch, fn, B = self.torch, self.fn, self.B
        nnz = fn.nnz_out(i)
        env_idx = torch.tensor(range(B), device=self.dev).repeat_interleave(nnz)
        row_idx = self._trip[i][0].to(self.dev).repeat(B)
        col_idx = self._trip[i][1].to(self.dev).repeat(B)
        dim = (B, fn.size1_out(i), fn.size2_out(i))
        return torch.sparse_coo_tensor(torch.vstack((env_idx, row_idx, col_idx)), self.outputs_sparse[i].reshape(-1), dim).to_dense()


def compat_tick(former, solver, inputs, rot, calls_per_tick, B, dev):
    """mpc_controller_cusadi.py:97-203 restated for measurement (everything between the marshalled inputs and the wrench),
    with `calls_per_tick` solver calls (4 for a 5-iteration kernel, 20 for the one-iteration build of oracle/_ref)."""
    import torch
    former.evaluate(inputs)
    H = former.getDenseOutput(0); f = former.getDenseOutput(1).squeeze(-1); A = former.getDenseOutput(2)
    b = former.getDenseOutput(3).squeeze(-1); G = former.getDenseOutput(4); d = former.getDenseOutput(5).squeeze(-1)
    Hs, fs, As, bs, Gs, ds = former.outputs_sparse
    H_col = Hs.double().contiguous(); G_col = Gs.double().contiguous(); A_col = As.double().contiguous()
    f_col = fs.reshape(B, -1, 1).double().contiguous(); d_col = ds.reshape(B, -1, 1).double().contiguous()
    b_col = bs.reshape(B, -1, 1).double().contiguous()
    nz, num_eq, num_ineq = H.shape[1], A.shape[1], G.shape[1]
    x = torch.zeros(B, nz, device=dev, dtype=torch.float64)
    s = torch.maximum(d - (G @ x.unsqueeze(2)).squeeze(2), torch.ones_like(d))
    z = torch.ones(B, num_ineq, device=dev, dtype=torch.float64)
    y = torch.ones(B, num_eq, device=dev, dtype=torch.float64)
    for _ in range(calls_per_tick):
        solver.evaluate([H_col, G_col, A_col, f_col, d_col, b_col, x.double().reshape(B, -1, 1).contiguous(),
                         s.double().reshape(B, -1, 1).contiguous(), z.double().reshape(B, -1, 1).contiguous(),
                         y.double().reshape(B, -1, 1).contiguous()])
        x = solver.getDenseOutput(0).reshape(B, -1).clone().detach(); s = solver.getDenseOutput(1).reshape(B, -1).clone().detach()
        z = solver.getDenseOutput(2).reshape(B, -1).clone().detach(); y = solver.getDenseOutput(3).reshape(B, -1).clone().detach()
    sol = x.clone().detach()
    u = sol[:, 120:132]
    lf, rf, lm, rm = u[:, :3].float(), u[:, 3:6].float(), u[:, 6:9].float(), u[:, 9:12].float()
    lm[:, 0] = 0.0; rm[:, 0] = 0.0
    Rt = rot.transpose(1, 2)
    parts = [-(Rt @ v.unsqueeze(-1)).squeeze(-1) for v in (lf, lm, rf, rm)]
    return torch.cat(parts, dim=1).reshape(-1, 2, 6).float()


def run_compat(args, emit=True, ticks=5):
    """`--impl compat`: what a maintainer gets who ONLY swaps the two shared objects and leaves
    CusadiFunction and MPCControllerCusadi.run() untouched: the reference tick (form + 4 x 5 iterations) at 4 096 envs
    through this repo's libqp_former.so / libsparse_pdipm_multiple_iterations.so, called exactly as the unmodified host
    code calls them -- next to the same host code over the reference's own generated kernels (oracle/_ref), and next to
    this repo's fused tick (srbd_mpc_solve, 20 iterations)."""
    import torch
    import biped_pympc_b200 as bp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cusadi_ref
    dev, B = "cuda:0", args.envs
    inp, rot = bp.workload.make_batch(B, seed=20261018)
    ins = [torch.as_tensor(a, device=dev) for a in inp]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=dev)
    # work tensor an unmodified CusadiFunction allocates and zero-fills: B x fn.sz_w() doubles; sz_w of the reference's
    # functions as built from its sources (oracle/_ref/*.meta), else the figures DESIGN.md section 1 records
    meta = {}
    for name, suffix, default in (("qp_former", "", 2179), ("sparse_pdipm_multiple_iterations", "", 7763)):
        path = os.path.join(cusadi_ref.REF_DIR, f"{name}{suffix}.meta")
        try:
            meta[name] = int(dict(line.split() for line in open(path))["sz_w"])
        except Exception:
            meta[name] = default
    libdir = os.path.dirname(bp._native.LIB_PATH)

    def measure(former, solver, calls):
        w = None
        for _ in range(2):
            w = compat_tick(former, solver, ins, rot_t, calls, B, dev)
        torch.cuda.synchronize()
        former.kernel_s = solver.kernel_s = 0.0
        t0 = time.perf_counter()
        for _ in range(ticks):
            w = compat_tick(former, solver, ins, rot_t, calls, B, dev)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / ticks
        return ms, 1e3 * former.kernel_s / ticks, 1e3 * solver.kernel_s / ticks, w

    ours_f = _CompatFunction(bp.QP_FORMER_FUNCTION, B, libdir, meta["qp_former"], dev)
    ours_s = _CompatFunction(bp.PDIPM_FUNCTION, B, libdir, meta["sparse_pdipm_multiple_iterations"], dev)
    ms, kf, ks, w = measure(ours_f, ours_s, 4)
    # the .so calls alone, same calling convention (zero-fill included), without the dense host-side work around them
    for o in (ours_f, ours_s):
        o.kernel_s = 0.0
    Hs, fs, As, bs, Gs, ds = ours_f.outputs_sparse
    x = torch.zeros(B, 240, dtype=torch.float64, device=dev); y = torch.ones(B, 140, dtype=torch.float64, device=dev)
    sv = torch.maximum(ds, torch.ones_like(ds)); z = torch.ones_like(ds)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ticks):
        ours_f.evaluate(ins)
        cur = [x, sv, z, y]
        for _ in range(4):
            ours_s.evaluate([Hs, Gs, As, fs, ds, bs, *cur])
            cur = [t.clone() for t in ours_s.outputs_sparse[:4]]
    torch.cuda.synchronize()
    calls_ms = 1e3 * (time.perf_counter() - t0) / ticks
    # this repo's fused tick on the same inputs: the wrench must agree (same 20 iterations)
    ctx = bp.Context(B, 0)
    wf = ctx.solve(ins, rot_t, num_calls=4, iters_per_call=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        wf = ctx.solve(ins, rot_t, num_calls=4, iters_per_call=5)
    torch.cuda.synchronize()
    fused_ms = 1e3 * (time.perf_counter() - t0) / 20
    agree = float((w - wf).abs().max() / wf.abs().max())
    ctx.close()
    rec = {"impl": "compat", "metric": "mpc_ticks_per_sec_4096_envs_4x5_iters_unmodified_host_code", "unit": "ticks/s (one env, one tick)",
           "value": B / (ms * 1e-3), "ms_per_tick": ms, "n_gpus": 1, "envs": B, "ticks_timed": ticks, "higher_is_better": True,
           "dtype": "f64", "data": "synthetic",
           "what": "the reference tick (qp_former + 4 x sparse_pdipm_multiple_iterations, mpc_controller_cusadi.py:97-203) through this "
                   "repo's two drop-in .so files, host code unmodified: zero-filled outputs and work tensor per call, 6 + 16 getDenseOutput "
                   "(sparse_coo -> dense), dense G @ x cold start, clones",
           "kernel_ms_per_tick": {"qp_former": kf, "sparse_pdipm_multiple_iterations_x4": ks},
           "evaluate_calls_only_ms_per_tick": calls_ms,
           "host_side_share": 1.0 - (kf + ks) / ms,
           "work_doubles_per_env_zero_filled": meta,
           "fused_tick_ms": fused_ms, "fused_tick_speedup_over_dropin": ms / fused_ms,
           "wrench_vs_fused_tick_rel": agree}
    var = cusadi_ref.solver_variant()
    if var is not None and cusadi_ref.available("qp_former"):
        suffix, per_call = var
        ref_f = _CompatFunction(bp.QP_FORMER_FUNCTION, B, cusadi_ref.REF_DIR, meta["qp_former"], dev)
        sz = int(dict(line.split() for line in open(os.path.join(cusadi_ref.REF_DIR, f"sparse_pdipm_multiple_iterations{suffix}.meta")))["sz_w"])
        ref_s = _CompatFunction(bp.PDIPM_FUNCTION, B, cusadi_ref.REF_DIR, sz, dev, suffix)
        ticks_ref, ticks = ticks, 2
        rms, rkf, rks, rw = measure(ref_f, ref_s, 20 // per_call)
        ticks = ticks_ref
        rec["reference_cusadi_same_host_code"] = {
            "ms_per_tick": rms, "kernel_ms_per_tick": {"qp_former": rkf, "solver": rks}, "solver_calls_per_tick": 20 // per_call,
            "what": "the same host code over the reference's own generated kernels (oracle/_ref)"
                    + ("" if per_call == 5 else "; solver = its function built with num_iters=1, 20 calls per tick"),
            "dropin_speedup": rms / ms, "wrench_vs_dropin_rel": float((rw - w).abs().max() / w.abs().max())}
    if emit:
        print(json.dumps(rec))
    return rec



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

    # ---- p50 / p99 step latency (the metric's second half; SURVEY 8d: 1 000 repetitions after 50 warm-ups, CUDA events
    #      around the stream work of one batched call, L2 flushed between repetitions) -------------------------------
    for _ in range(50):
        step()
    lat_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.latency_reps)]
    torch.cuda.synchronize()
    t_lat0 = time.time()
    for a, b in lat_ev:
        flush.fill_(1.0)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    t_lat1 = time.time()
    lat_ms = sorted(a.elapsed_time(b) for a, b in lat_ev)
    # clocks and throttle reasons: sampled over the K timed steps AND the latency loop (the K steps alone are shorter than
    # one 20 ms sampling period at the default K)
    clocks = sampler.stop(t_wall0, t_lat1)
    clocks["window"] = "timed steps + latency loop" if clocks.get("window") == "timed region" else clocks.get("window")

    # (the e2e leg runs AFTER the clock sampler has been stopped: `nvidia-smi -lms 20` polls the driver every 20 ms and
    #  contends with the ~20 CUDA API calls of a host-entry tick -- 0.71 ms in a quiet process, 0.77 ms next to the poller)
    # ---- e2e: host buffers through the C-ABI (H2D + kernels + D2H inside the timed region) ----
    # x, u (inputs 1, 2) are the linearisation point, all ones at the reference call site
    # (mpc_controller_cusadi.py:55-56) and dead for this LTV-linear problem: the host API takes
    # None for them and does not copy them
    assert all(float(inp[i].min()) == 1.0 == float(inp[i].max()) for i in (1, 2))
    h_in = [None if i in (1, 2) else torch.as_tensor(a).pin_memory() for i, a in enumerate(inp)]
    h_rot = torch.as_tensor(rot, dtype=torch.float32).pin_memory()
    h_w = torch.empty((B, 2, 6), dtype=torch.float32).pin_memory()
    # the buffers are bound once (validation + pointer marshalling of 19 tensors, ~45 us of Python per call otherwise); every
    # call below still moves every input byte host -> device and the wrench device -> host inside the timed region
    host_tick = ctx.bind_host(h_in, h_rot, h_w)
    # warm-up: the first call pays one-time set-up (1.1-1.4 ms), and for some tens of milliseconds after the clock sampler's
    # nvidia-smi process has been terminated the calls take 0.70-0.75 ms instead of 0.60 (measured per call); both are over
    # after 40 calls
    time.sleep(0.1)
    for _ in range(40):
        host_tick(args.calls, args.iters, args.tol)
    n_e2e = max(10, min(args.steps, 100))
    barrier()
    e2e_calls = []
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        t_c = time.perf_counter()
        host_tick(args.calls, args.iters, args.tol)   # synchronous
        e2e_calls.append(time.perf_counter() - t_c)
    e2e_s = time.perf_counter() - t0
    if os.environ.get("SRBD_BENCH_E2E_DEBUG"):
        sc = sorted(e2e_calls)
        sys.stderr.write("e2e calls ms: first " + " ".join("%.3f" % (1e3 * t) for t in e2e_calls[:6]) + " | median %.4f p90 %.4f max %.4f mean %.4f\n"
                         % (1e3 * sc[len(sc) // 2], 1e3 * sc[int(0.9 * len(sc))], 1e3 * sc[-1], 1e3 * e2e_s / n_e2e))
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
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
                "steps": n_e2e, "api": "Context.bind_host(...)() = srbd_mpc_solve_host on buffers validated once (pinned host buffers read and written over PCIe by the kernels themselves: the first 2368-env chunk -- one full round of the persistent kernel -- in place, the rest pulled into device staging by the copy engine while the first solves)"},
        "gpu_launches": args.steps,          # ONE kernel per step: fused assembly + PDIPM + wrench (pdipm_fast_kernel)
        "clocks": clocks,
    }
    ts = sorted(times_ms)
    pct = lambda v, q: v[min(len(v) - 1, int(q * len(v)))]  # noqa: E731
    line["latency_ms"] = {"p50": pct(lat_ms, 0.50), "p99": pct(lat_ms, 0.99), "max": lat_ms[-1], "repetitions": len(lat_ms),
                          "what": "one batched call (CUDA events on its stream), separate loop after 50 warm-ups, L2 flushed between calls",
                          "timed_steps": {"p50": pct(ts, 0.50), "p99": pct(ts, 0.99), "max": ts[-1], "n": len(ts)}}

    if iters_done is not None:
        hist = torch.bincount(iters_done.clamp(min=0).to(torch.int64), minlength=args.calls * args.iters + 1).tolist()
        line["iterations_histogram"] = {str(i): c for i, c in enumerate(hist) if c}
    n_it = args.calls * args.iters
    subs = default_wl and not args.no_sub
    if subs:                      # every rank takes part (65 536 / N envs each)
        line["strong_scaling_65536"] = sub_strong_scaling(bp, torch, dist, dev, local_rank, rank, world, flush)
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
        asm_small = {"kernel": "qp_form_kernel", "envs": B, "bound": "launch latency + L2 (the 88.7 MB of output stay in the 126 MB L2: not an HBM-bound launch)",
                     "achieved": ach_bw, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": ach_bw / peaks.get("hbm_gbs", 6650.0),
                     "traffic": cap.get("assembly_dram_bytes_per_launch"),
                     "traffic_note": "DRAM bytes of one launch from profiles/ncu_qpform_r1_4096_summary.csv: below the "
                                     f"{B * BYTES_ASSEMBLY} algorithmic bytes because most of the output is still in L2 at kernel end",
                     "kernel_ms": fm_ms, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"}
        if subs:
            # the assembly kernel's HBM roofline is quoted where the launch IS HBM-bound: 65 536 envs (1.4 GB per launch)
            B65 = 65536
            ins65 = [t.repeat((B65 + B - 1) // B, 1)[:B65].contiguous() for t in ins]
            qp65 = bp.qp_form(ins65)
            f65 = []
            for i in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                bp.qp_form(ins65, out=qp65)
                b.record()
                b.synchronize()
                if i >= 3:
                    f65.append(a.elapsed_time(b))
            ms65 = sum(f65) / len(f65)
            bw65 = B65 * BYTES_ASSEMBLY / (ms65 * 1e-3) / 1e9
            line["roofline_assembly"] = {"kernel": "qp_form_kernel", "envs": B65, "bound": "hbm", "achieved": bw65, "peak": peaks.get("hbm_gbs"),
                                         "unit": "GB/s", "frac": bw65 / peaks.get("hbm_gbs", 6650.0), "traffic": None, "kernel_ms": ms65,
                                         "algorithmic": f"{BYTES_ASSEMBLY} B/env x {B65} envs (inputs and outputs larger than L2: no flush needed)",
                                         "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)", "at_metric_batch": asm_small}
            del ins65, qp65
        else:
            line["roofline_assembly"] = asm_small
        # ---- CPU baseline on the box's host cores (bounded sample) -----------------------------
        if world == 1 and default_wl:
            threads = os.cpu_count() or 1
            cpu_val, _, kind, desc = cpu_reference_run(B, ITERS, 3, 1, threads)
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"the metric's full batch, {B} envs x 3 steps of the same seeded workload; {desc}"}
            port_val, _, _, port_desc = cpu_reference_run(B, ITERS, 3, 1, threads, want="port")
            line["cpu_baseline_port"] = {"value": port_val, "unit": UNIT, "cores": threads, "kind": "port",
                                         "sample": f"{B} envs x 3 steps; {port_desc}"}
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
        if world == 1 and subs:
            # ---- the e2e leg's other cases: every one of the 17 buffers passed (x, u included: byte for byte what the
            #      reference hands to qp_former), pinned and pageable (cudaMemcpyAsync staging path) ----------------------
            def e2e_case(pin):
                mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
                hi = [mk(torch.as_tensor(a).clone()) for a in inp]
                hr, hw = mk(torch.as_tensor(rot, dtype=torch.float32).clone()), mk(torch.empty((B, 2, 6), dtype=torch.float32))
                for _ in range(3):
                    ctx.solve_host(hi, hr, hw, num_calls=args.calls, iters_per_call=args.iters)
                t0 = time.perf_counter()
                for _ in range(20):
                    ctx.solve_host(hi, hr, hw, num_calls=args.calls, iters_per_call=args.iters)
                dt = (time.perf_counter() - t0) / 20
                return {"value": B / dt, "unit": UNIT, "ms_per_step": 1e3 * dt, "steps": 20,
                        "h2d_bytes_per_step": sum(a.numel() * 8 for a in hi) + hr.numel() * 4, "d2h_bytes_per_step": hw.numel() * 4}
            line["e2e_all17_pinned"] = e2e_case(True)
            line["e2e_all17_pageable"] = e2e_case(False)
            for key, fn in (("single_env_config0", lambda: sub_single_env(bp, torch, dev, local_rank)),
                            ("fp32_mode", lambda: sub_fp32_mode(bp, torch, dev, local_rank, flush, line["ms_per_step"])),
                            ("converge_16384", lambda: sub_converge(bp, torch, dev, local_rank, flush)),
                            ("closed_loop_4096x1000", sub_closed_loop),
                            ("compat_dropin", lambda: run_compat(argparse.Namespace(envs=B), emit=False))):
                try:
                    line[key] = fn()
                except Exception as exc:      # a sub-record must not take the bench line down
                    line[key] = {"unavailable": repr(exc)}
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cusadi", "compat"])
    ap.add_argument("--latency-reps", type=int, default=1000, help="repetitions of the p50 / p99 latency loop")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (other BASELINE configs, drop-in, closed loop)")
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
    elif args.impl == "compat":
        if int(os.environ.get("RANK", "0")) == 0:
            run_compat(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
