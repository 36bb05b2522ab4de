"""Developer tool: the closed-loop rollout of tests/closed_loop.py on the CUDA path, cold- and warm-started."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import closed_loop as cl, biped_pympc_b200 as bp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for mode, iters in (("cold", 20), ("cold", 30), ("warm", 10), ("warm", 15), ("warm", 20)):
    ro = cl.Rollout(B, seed=11, walking=True)
    ctx = bp.Context(B)
    rot = torch.eye(3, dtype=torch.float32, device="cuda:0")[None].repeat(B, 1, 1)
    tsolve = [0.0]

    def solve(ins):
        dev = [torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") for a in ins]
        torch.cuda.synchronize(); t = time.time()
        if mode == "cold":
            x = ctx.solve(dev, rot, num_calls=1, iters_per_call=iters, want_iterates=True)[1]
        else:
            x = ctx.solve_warm(dev, rot, iters=iters, want_iterates=True)[1]
        torch.cuda.synchronize(); tsolve[0] += time.time() - t
        return x.cpu().numpy()
    ro.run(solve, ticks)
    herr = np.abs(ro.p[:, 2] - cl.HZ)
    s = ro.summary()
    print(f"{mode:4s} {iters:2d} it/tick: envs with height error > 8 cm: {int((herr > 0.08).sum()):4d} of {B} | "
          f"height_err_max {s['height_err_max']:.3f} tilt_max {s['tilt_max']:.3f} vx_err_mean {s['vx_err_mean']:.3f} "
          f"cone {s['cone_violation_max']:.1e} finite {s['finite']} | {1e3 * tsolve[0] / ticks:.2f} ms/tick", flush=True)
    ctx.close()
