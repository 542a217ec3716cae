"""Test-side "local solver" for swe_fvm_b200.dist: the CPU oracle behind the same small interface
as GpuLocal, so the decomposition / halo / dt-reduction host logic runs under gloo without a GPU."""
import numpy as np
import torch


class OracleLocal:
    def __init__(self, oracle, flux=1, ws=2):
        self.o, self.flux, self.ws = oracle, flux, ws
        self.dt = 0.0
        self._ml = torch.ones(1, dtype=torch.float64)
        self.U0 = None
        self.send = self.recv = None

    def alloc(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def set_halo_lists(self, send, recv):
        self.send, self.recv = np.asarray(send, dtype=np.int64), np.asarray(recv, dtype=np.int64)

    def set_cfl_edge_mask(self, mask):
        self.o.set_cfl_edge_mask(mask)

    def pack(self, buf):
        st = self.o.get_state()
        buf[:3 * len(self.send)] = torch.from_numpy(st[self.send].reshape(-1))

    def unpack(self, buf):
        st = self.o.get_state()
        st[self.recv] = buf[:3 * len(self.recv)].numpy().reshape(-1, 3)
        self.o.set_state(st)

    def min_len_tensor(self):
        self._ml[0] = self.o.min_len_to_wavespeed()
        return self._ml

    def compute_interface_values(self):
        self.o.compute_interface_values()

    def compute_fluxes(self):
        self.o.compute_fluxes(self.flux, self.ws)

    def save_state(self):
        self.U0 = self.o.get_state()

    def stage_update(self, a0, a1, coef, dt):
        dts = coef * (self.dt if dt is None else dt)
        if a0 == 0.0:
            self.o.stage_update(None, 0.0, 1.0, dts, True)
        else:
            self.o.stage_update(self.U0, a0, a1, dts, False)

    def set_dt(self, dt):
        self.dt = float(dt)

    def advance_dt(self, dt):
        if dt is None:
            self.dt = 0.15 * float(self._ml[0])
