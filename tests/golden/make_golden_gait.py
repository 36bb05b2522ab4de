"""Generates tests/golden/srbd_ref_gait_v1.npz by running the reference's GaitGenerator.

biped_pympc/core/gait/gait_generator.py is pure torch: it is imported unmodified from /root/reference
(without the package __init__, which pulls in CasADi) and stepped on the CPU the way BipedController
does (biped_controller.py:95-127, 131-137): every control step get_swing_sub_phase, get_contact_sub_phase,
mpc_gait, then update_phase; the MPC sampling time is changed mid-run (update_sampling_time, as RL-MPC does:
mpc_wrapper.py:48-56).  Gaits: the default walking gait (ssp 5, dsp 0: configuration.py:17-18), a walking
gait with double support, an asymmetric one, and the standing gait (biped_controller.py:26-28), with the
phase started at random values.
Run:  python tests/golden/make_golden_gait.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

pkg = types.ModuleType("biped_pympc")
pkg.__path__ = [os.path.join(REF, "biped_pympc")]
sys.modules["biped_pympc"] = pkg
from biped_pympc.core.gait.gait_generator import GaitGenerator  # noqa: E402

GAITS = [((5, 5), (0, 0)), ((5, 5), (2, 2)), ((4, 6), (1, 1)), ((0, 0), (5, 5))]     # (ssp, dsp)
PER, STEPS, HORIZON, DT = 8, 400, 10, 0.001


def main():
    g = torch.Generator().manual_seed(20261018)
    ssp = torch.tensor([s for s, d in GAITS for _ in range(PER)])
    dsp = torch.tensor([d for s, d in GAITS for _ in range(PER)])
    B = ssp.shape[0]
    dt_mpc = 0.02 + 0.03 * torch.rand(B, generator=g)
    gen = GaitGenerator(B, HORIZON, DT, dt_mpc.clone(), dsp, ssp)
    gen.gait_phase[:] = torch.rand(B, generator=g)
    gen.gait_phase[0] = 0.0
    out = {"ssp": ssp.numpy().astype(np.int32), "dsp": dsp.numpy().astype(np.int32), "dt": np.array(DT),
           "horizon": np.array(HORIZON), "dt_mpc0": dt_mpc.numpy().copy(), "phase0": gen.gait_phase.numpy().copy()}
    dt_mpc1 = 0.02 + 0.03 * torch.rand(B, generator=g)
    out["dt_mpc1"] = dt_mpc1.numpy().copy()
    out["switch_step"] = np.array(STEPS // 2)
    rec = {k: [] for k in ("phase", "contact_sub", "swing_sub", "table")}
    for t in range(STEPS):
        if t == STEPS // 2:
            gen.update_sampling_time(dt_mpc1.clone())
        rec["phase"].append(gen.gait_phase.numpy().copy())
        rec["swing_sub"].append(gen.get_swing_sub_phase().numpy().copy())
        rec["contact_sub"].append(gen.get_contact_sub_phase().numpy().copy())
        rec["table"].append(gen.mpc_gait.numpy().copy().astype(np.int8))
        gen.update_phase()
    rec["phase"].append(gen.gait_phase.numpy().copy())
    for k, v in rec.items():
        out[k] = np.stack(v)
    out["swing_durations_sec"] = gen.swing_durations_sec.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "srbd_ref_gait_v1.npz"), **out)
    print("saved srbd_ref_gait_v1.npz", B, "envs", STEPS, "steps; wraps:", int((np.diff(out["phase"], axis=0) < 0).sum()))


if __name__ == "__main__":
    main()
