"""Generates tests/golden/srbd_ref_controller_v1.npz by running the reference's OWN controller.

`MPCControllerCusadi.run()` (biped_pympc/convex_mpc/mpc_controller_cusadi.py:43-205) and its base class
(convex_mpc/base_controller.py:61-257) are imported unmodified from /root/reference and executed on the
CPU for two consecutive MPC ticks of a seeded batch: knot points, horizon state, initial state,
reference trajectory, the marshalling into the 17 qp_former inputs (with its layout quirks Q1-Q3,
SURVEY 8b), the cold start, the four chained solver calls and the wrench post-processing are all the
reference's code.  Two things the container does not have are stood in for:

  * `casadi`            -> tests/golden/casadi_shim/casadi.py (see make_golden_ref.py); the two
                           serialized functions the controller loads (`srbd_qp_mat.casadi`,
                           `mpc_multiple_iter_5_solver_240v_140eq_160ineq.casadi`, absent from the
                           repository: script/setup.sh:3-13) are rebuilt from the reference's
                           srbd_constraints.py / sparse_pdipm_solver.py as generate_solver_function.py does
  * `CusadiFunction`    -> `CpuCusadiFunction` below: same public surface (evaluate / outputs_sparse /
                           getDenseOutput, CusadiFunction.py:28-58) but evaluates the CasADi function per
                           env on the CPU instead of launching the generated CUDA kernel.  It reads the
                           caller's tensors exactly like that kernel does -- env `idx` takes
                           flat[idx*nnz_in : (idx+1)*nnz_in] of the contiguous buffer
                           (generateCUDACode.py:126-128) -- which is what produces the quirks.

Recorded per tick: controller state in, x0 / x_ref / contact table as the base class built them, the 17
tensors handed to qp_former, the final iterate and the returned foot wrench.
Run:  python tests/golden/make_golden_controller.py        (about a minute)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_ref as mgr  # noqa: E402

B = 12


class CpuCusadiFunction:
    def __init__(self, fn_casadi, num_instances):
        self.fn_casadi = fn_casadi
        self.fn_name = fn_casadi.name()
        self.num_instances = num_instances
        self.outputs_sparse = [torch.zeros(num_instances, fn_casadi.nnz_out(i), dtype=torch.double)
                               for i in range(fn_casadi.n_out())]
        self.inputs_sparse = []
        self.calls = []

    def evaluate(self, inputs):
        fn = self.fn_casadi
        assert len(inputs) == fn.n_in()
        self.inputs_sparse = inputs
        self.calls.append([t.detach().clone() for t in inputs])
        for o in self.outputs_sparse:
            o.zero_()
        flats = []
        for i, t in enumerate(inputs):
            assert t.dtype == torch.double and t.is_contiguous()
            flats.append(t.reshape(-1).numpy())
        for idx in range(self.num_instances):
            args = [f[idx * fn.nnz_in(i):(idx + 1) * fn.nnz_in(i)] for i, f in enumerate(flats)]
            res = fn(*args)
            res = res if isinstance(res, tuple) else (res,)
            for o, r in zip(self.outputs_sparse, res):
                o[idx] = torch.tensor(r.nonzeros(), dtype=torch.double)

    def getDenseOutput(self, out_idx=None):
        sp = self.fn_casadi.sparsity_out(out_idx)
        colptr, row = sp.get_ccs()
        rows, cols = sp.mask.shape
        dense = torch.zeros(self.num_instances, rows, cols, dtype=torch.double)
        col = np.repeat(np.arange(cols), np.diff(colptr))
        dense[:, torch.as_tensor(row), torch.as_tensor(col)] = self.outputs_sparse[out_idx]
        return dense


def main():
    SRBD_Constraints, sparse_pdipm_multiple_iterations = mgr.import_reference()
    import casadi
    wl = mgr.load_workload()

    # the two functions the controller loads, rebuilt as generate_solver_function.py:60-114 does
    qp_former = SRBD_Constraints(10).qp_mat_former
    H, f, A, b, G, d = qp_former(*[a[0] for a in wl.nominal_standing_inputs()])
    (Hc, Hr), (Ac, Ar), (Gc, Gr) = H.sparsity().get_ccs(), A.sparsity().get_ccs(), G.sparsity().get_ccs()
    solver = sparse_pdipm_multiple_iterations(240, 140, 160, Hr, Hc, Ar, Ac, Gr, Gc, beta=1e-8, num_iters=5)
    table = {"srbd_qp_mat.casadi": qp_former, "mpc_multiple_iter_5_solver_240v_140eq_160ineq.casadi": solver}
    casadi.Function.load = staticmethod(lambda path: table[os.path.basename(path)])

    fake = types.ModuleType("biped_pympc.cusadi")
    fake.CASADI_FUNCTION_DIR = "/nonexistent/casadi_functions"
    fake.CusadiFunction = CpuCusadiFunction
    sys.modules["biped_pympc.cusadi"] = fake

    from biped_pympc.configuration.configuration import MPCConf
    from biped_pympc.convex_mpc.mpc_controller_cusadi import MPCControllerCusadi
    from biped_pympc.core.data.robot_data import DesiredStateData, StateEStimatorData

    cfg = MPCConf(print_solve_time=False)
    cfg.Q = cfg.Q.cpu()
    cfg.R = cfg.R.cpu()
    ctrl = MPCControllerCusadi(B, "cpu", 2, cfg)
    out = {"mass": np.array(float(ctrl.mass)), "mu": np.array(float(ctrl.mu)), "I_body": ctrl.biped.I_body.numpy(),
           "Q": cfg.Q.numpy(), "R": cfg.R.numpy(), "dt": np.array(cfg.dt), "decimation": np.array(cfg.decimation)}

    g = torch.Generator().manual_seed(20261018)
    rnd = lambda *s: torch.randn(*s, generator=g)      # noqa: E731
    uni = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, generator=g)  # noqa: E731
    sed = StateEStimatorData(2, B, torch.device("cpu"))
    des = DesiredStateData(B, torch.device("cpu"))
    euler = torch.stack([0.1 * rnd(B), 0.1 * rnd(B), uni(-3.1, 3.1, B)], 1)
    pos = torch.stack([rnd(B), rnd(B), 0.55 + 0.02 * rnd(B)], 1)
    vel_cmd = uni(-0.5, 0.5, B, 2)
    vel_cmd[:3, 0] = torch.tensor([0.0, 0.005, -0.009])          # stationary_mask branch (|vx| < 1e-2)
    yaw_rate = uni(-0.5, 0.5, B)
    des.set_command(vel_cmd, yaw_rate, 0.55 + 0.01 * rnd(B))
    dt_mpc = uni(0.02, 0.05, B)
    phase = uni(0.0, 1.0, B).numpy()
    for tick in range(2):
        rot = torch.as_tensor(wl._rot_zyx(euler[:, 0].numpy().astype(np.float64), euler[:, 1].numpy().astype(np.float64),
                                          euler[:, 2].numpy().astype(np.float64)), dtype=torch.float32)
        sed.root_euler[:] = euler
        sed.rotation_body[:] = rot
        sed.root_position[:] = pos
        sed.root_angular_velocity_w[:] = 0.2 * rnd(B, 3)
        sed.root_velocity_w[:] = 0.3 * rnd(B, 3)
        hip = torch.zeros(B, 2, 3)
        hip[:, 0, 1], hip[:, 1, 1] = 0.106, -0.106
        feet = pos[:, None, :] + torch.einsum("bij,bfj->bfi", rot, hip)
        feet[:, :, :2] += 0.03 * rnd(B, 2, 2)
        feet[:, :, 2] = 0.0
        sed.foot_position[:] = feet
        contact = torch.as_tensor(wl.mpc_gait_table((phase + 0.1 * tick) % 1.0), dtype=torch.float32)
        ctrl.set_state_estimate_data(sed)
        ctrl.set_desired_state_data(des)
        ctrl.set_contact_table(contact)
        ctrl.set_mpc_sampling_time(dt_mpc)
        ctrl.residual_lin_accel[:] = 0.5 * rnd(B, 3)
        ctrl.residual_ang_accel[:] = 0.5 * rnd(B, 3)
        t = "t%d_" % tick
        for name in ("root_euler", "rotation_body", "root_position", "root_angular_velocity_w", "root_velocity_w",
                     "foot_position"):
            out[t + name] = getattr(sed, name).numpy().copy()
        out[t + "desired_velocity_b"] = des.desired_velocity_b.numpy().copy()
        out[t + "desired_angular_velocity_b"] = des.desired_angular_velocity_b.numpy().copy()
        out[t + "desired_height"] = des.desired_height.numpy().copy()
        out[t + "contact_table_in"] = contact.numpy().copy()
        out[t + "dt_mpc"] = dt_mpc.numpy().copy()
        out[t + "residual_lin_accel"] = ctrl.residual_lin_accel.numpy().copy()
        out[t + "residual_ang_accel"] = ctrl.residual_ang_accel.numpy().copy()
        out[t + "world_position_desired_in"] = ctrl.world_position_desired.numpy().copy()
        out[t + "yaw_desired_in"] = ctrl.yaw_desired.numpy().copy()
        out[t + "first_run_in"] = ctrl.first_run.numpy().copy()

        wrench, cost = ctrl.run()

        out[t + "x0"] = ctrl.x0.numpy().copy()
        out[t + "x_ref"] = ctrl.x_ref.numpy().copy()
        out[t + "world_position_desired_out"] = ctrl.world_position_desired.numpy().copy()
        out[t + "yaw_desired_out"] = ctrl.yaw_desired.numpy().copy()
        for n, a in zip(wl.INPUT_NAMES, ctrl.qp_former.calls[-1]):
            out[t + "in17_" + n] = a.numpy().copy()
        assert len(ctrl.qp_solver.calls) == 4 * (tick + 1)
        for n, a in zip(("x", "s", "z", "y", "res", "mu"), ctrl.qp_solver.outputs_sparse):
            out[t + "ip20_" + n] = a.numpy().copy()
        out[t + "wrench"] = wrench.numpy().copy()
        out[t + "cost"] = cost.numpy().copy()
        print("tick", tick, "wrench[0]", wrench[0].numpy().round(2).tolist(), flush=True)
        # next tick: the body has moved a little
        pos = pos + 0.01 * torch.cat([vel_cmd, torch.zeros(B, 1)], 1)
        euler = euler + torch.stack([torch.zeros(B), torch.zeros(B), 0.01 * yaw_rate], 1)
    np.savez_compressed(os.path.join(HERE, "srbd_ref_controller_v1.npz"), **out)
    print("saved srbd_ref_controller_v1.npz")


if __name__ == "__main__":
    main()
