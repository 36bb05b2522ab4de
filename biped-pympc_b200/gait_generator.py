"""Device-side counterpart of ``GaitGenerator`` (biped_pympc/core/gait/gait_generator.py:3-252).

Same constructor arguments, attributes and methods; every query is one launch of ``srbd_gait`` instead of
a chain of small torch kernels with boolean-mask indexing (which synchronises the host), and ``tick()``
returns everything a control step needs -- both sub-phases, the MPC contact table -- and advances the
phase, in ONE launch.  fp32 arithmetic in the reference's operation order: results are bit-identical to
the reference's (tests/test_gpu_parity.py::test_gait_kernel_matches_the_reference_gait_generator, against a
recording of the reference class).  There is no CPU path: the tensors must live on a CUDA device.
"""
from __future__ import annotations

from . import _native


class GaitGeneratorB200:
    def __init__(self, batch_size: int, mpc_horizon: int, dt: float, dt_mpc, dsp_durations, ssp_durations, device="cuda"):
        import torch
        self.batch_size, self.mpc_horizon, self.dt, self.device = batch_size, mpc_horizon, float(dt), torch.device(device)
        if self.device.type != "cuda":
            raise _native.SrbdError("GaitGeneratorB200 needs a CUDA device (no CPU fallback)")
        _native.load()
        i32 = lambda t: torch.as_tensor(t).to(torch.int32).to(self.device).contiguous()  # noqa: E731
        self.dsp_durations, self.ssp_durations = i32(dsp_durations), i32(ssp_durations)      # :48-49
        self.gait_cycle_length = torch.sum(self.dsp_durations + self.ssp_durations, dim=1)
        self.gait_phase = torch.zeros(batch_size, device=self.device)
        self.stance_durations = torch.stack([self.ssp_durations[:, 1] + torch.sum(self.dsp_durations, dim=1),
                                             self.ssp_durations[:, 0] + torch.sum(self.dsp_durations, dim=1)], dim=1)
        self.swing_durations = torch.stack([self.ssp_durations[:, 0], self.ssp_durations[:, 1]], dim=1)
        self.ssp_durations_phase = self.ssp_durations / self.gait_cycle_length.unsqueeze(1)
        self.dsp_durations_phase = self.dsp_durations / self.gait_cycle_length.unsqueeze(1)
        self.mpc_table = torch.ones((batch_size, mpc_horizon, 2), dtype=torch.int32, device=self.device)
        self.contact_table = torch.ones((batch_size, mpc_horizon, 2), dtype=torch.float32, device=self.device)
        self.contact_sub_phase = -torch.ones((batch_size, 2), device=self.device)
        self.swing_sub_phase = -torch.ones((batch_size, 2), device=self.device)
        self.update_sampling_time(dt_mpc)

    def reset(self, env_ids=None):
        self.gait_phase[env_ids] = 0.0

    def update_sampling_time(self, dt_mpc):
        import torch
        self.dt_mpc = torch.as_tensor(dt_mpc, dtype=torch.float32, device=self.device).contiguous()
        self.swing_durations_sec = self.swing_durations * self.dt_mpc[:, None]
        self.stance_durations_sec = self.stance_durations * self.dt_mpc[:, None]
        self.gait_durations_sec = self.gait_cycle_length * self.dt_mpc

    def _launch(self, advance, contact=None, swing=None, table_i32=None, table_f32=None):
        import torch
        lib = _native.load()
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(self.device):
            _native.check(lib.srbd_gait(self.gait_phase.data_ptr(), self.dt_mpc.data_ptr(), self.ssp_durations.data_ptr(),
                                        self.dsp_durations.data_ptr(), self.dt, int(advance), self.mpc_horizon, p(contact),
                                        p(swing), p(table_i32), p(table_f32), self.batch_size, _native._stream()))

    def update_phase(self):
        self._launch(True)

    def get_contact_sub_phase(self):
        self._launch(False, contact=self.contact_sub_phase)
        return self.contact_sub_phase

    def get_swing_sub_phase(self):
        self._launch(False, swing=self.swing_sub_phase)
        return self.swing_sub_phase

    @property
    def contact_bool(self):
        return self.get_contact_sub_phase() != -1

    @property
    def swing_bool(self):
        return self.get_swing_sub_phase() != -1

    @property
    def mpc_gait(self):
        self._launch(False, table_i32=self.mpc_table, table_f32=self.contact_table)
        return self.mpc_table

    def tick(self, advance: bool = True):
        """One launch per control step: (contact_sub_phase, swing_sub_phase, contact_table fp32 (B,H,2)) of the
        current phase, then update_phase -- the order of BipedController.run_lowlevel (biped_controller.py:120-127)."""
        self._launch(advance, self.contact_sub_phase, self.swing_sub_phase, self.mpc_table, self.contact_table)
        return self.contact_sub_phase, self.swing_sub_phase, self.contact_table
