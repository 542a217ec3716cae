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


class EmulatedSolver:
    """Test-side stage loop of the multi-rank step over torch.distributed (gloo) with any local solver
    (here the CPU oracle): the host logic of csrc/swe_dist.cuh restated in Python — one halo exchange
    per stage with the lists of the C++ plan, one min all-reduce per step. It checks that the PLAN
    (owned / halo / send / receive lists, CFL edge mask) gives bit-identical owned cells."""

    STAGES = {0: [(0.0, 1.0, 1.0)], 1: [(0.0, 1.0, 1.0), (0.5, 0.5, 0.5)],
              2: [(0.0, 1.0, 1.0), (0.75, 0.25, 0.25), (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0)]}

    def __init__(self, plan, local):
        self.plan, self.local = plan, local
        self.nsend = sum(len(p[1]) for p in plan.peers)
        self.nrecv = sum(len(p[2]) for p in plan.peers)
        local.set_halo_lists(plan.send_list(), plan.recv_list())
        local.set_cfl_edge_mask(plan.cfl_mask)
        self.sendbuf = local.alloc(3 * max(self.nsend, 1))
        self.recvbuf = local.alloc(3 * max(self.nrecv, 1))
        self.exchanges = 0

    def exchange(self):
        import torch.distributed as dist
        if not self.plan.peers:
            return
        self.local.pack(self.sendbuf)
        ops, so, ro = [], 0, 0
        for peer, s, r in self.plan.peers:
            if len(s):
                ops.append(dist.P2POp(dist.isend, self.sendbuf[3 * so:3 * (so + len(s))], peer))
            if len(r):
                ops.append(dist.P2POp(dist.irecv, self.recvbuf[3 * ro:3 * (ro + len(r))], peer))
            so += len(s)
            ro += len(r)
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.local.unpack(self.recvbuf)
        self.exchanges += 1

    def step(self, scheme, dt):
        import torch.distributed as dist
        L = self.local
        stages = self.STAGES[scheme]
        for k, (a0, a1, coef) in enumerate(stages):
            L.compute_interface_values()
            L.compute_fluxes()
            if k == len(stages) - 1 and self.plan.world > 1:
                dist.all_reduce(L.min_len_tensor(), op=dist.ReduceOp.MIN)
            if k == 0 and len(stages) > 1:
                L.save_state()
            L.stage_update(a0, a1, coef, dt)
            self.exchange()
        L.advance_dt(dt)

    def run(self, scheme, nsteps, dt, dt0=0.0):
        if dt is None:
            self.local.set_dt(dt0)
        for _ in range(nsteps):
            self.step(scheme, dt)
