"""Host layer of K2: one huge lattice updated by the sublattice-parallel kernel (BASELINE config 5).

The lattice is a byte array in HBM (0 empty / 1 '+' / 2 '-'), ping-ponged between two buffers.  With
more than one rank the lattice is cut into contiguous slabs (one per GPU) that carry a ghost zone of
`ghost` sites on each interior side.  Because the random stream is keyed by the GLOBAL segment index,
a rank recomputes its ghost zone bit-identically to its neighbour; the ghost data goes stale from its
outer end inwards by at most (radius + 33) sites per pass — the trials of an active half of 32 sites run
sequentially, so one stale site can change a chain of hops across the whole half, plus the reach of the
field — so the ghosts are refreshed every `(ghost - radius - 1) // (radius + 33)` passes, not every pass.

CUDA backend: `run_passes(n)` is ONE cooperative launch (`aps_k2_run_persistent_device`): grid barriers
between the passes, and with several ranks the kernel itself refreshes the ghost zones and (global-field mode)
exchanges the 8-byte sum(sigma) increments through peer memory mapped with CUDA IPC over NVLink — flag words
written with system-scope release stores, no NCCL call and no host round trip inside the time stepping
(csrc/aps_k2.cuh).  torch.distributed is only used once, to hand the 64-byte IPC handles round, and for the
observables (profile all-reduce).  The oracle backend of the CPU tests keeps the host-driven exchange below.
The result is bit-identical to the single-slab run (gloo tests on CPU with the oracle backend, pytest -m gpu on
2 GPUs with the kernels).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import ApsK2Args, ApsK2Rates
from .engine import gaussian_weights
from .launcher import dist_info

TILE = 8192


def fixed_point_taps(sigma_sites: float):
    """Gaussian taps in lattice units as 2^-16 fixed point, centre first (aps_k2_model.h)."""
    r, w = gaussian_weights(sigma_sites)
    return r, np.ascontiguousarray(np.round(w[r:] * 65536.0).astype(np.int32))


class CudaK2Backend:
    """Product backend: buffers are torch CUDA tensors, passes go through the C ABI."""
    name = "cuda"

    def __init__(self):
        self.lib = capi.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())

    def zeros_u8(self, n):
        return torch.zeros(n, dtype=torch.uint8, device=self.dev)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def zeros_i64(self, n):
        return torch.zeros(n, dtype=torch.int64, device=self.dev)

    def ptr(self, t):
        return t.data_ptr()

    def rates(self, D, lam, beta, dt):
        r = ApsK2Rates()
        capi.check(self.lib.aps_k2_rates_init(D, lam, beta, dt, r), "aps_k2_rates_init")
        return r

    def flip_table(self, rates):
        t = np.zeros(2 * 1025, np.uint32)
        capi.check(self.lib.aps_k2_flip_table(rates, t.ctypes.data), "aps_k2_flip_table")
        return torch.from_numpy(t.view(np.int32).copy()).to(self.dev)

    def run(self, args, n_passes):
        capi.check(self.lib.aps_k2_run_device(args, n_passes, torch.cuda.current_stream().cuda_stream), "aps_k2_run_device")

    persistent = True          # run_passes() = one cooperative multi-pass launch with in-kernel exchange

    def run_persistent(self, args, multi):
        capi.check(self.lib.aps_k2_run_persistent_device(args, multi, torch.cuda.current_stream().cuda_stream),
                   "aps_k2_run_persistent_device")

    def peer_setup(self, rank, world):
        """Allocate this rank's exchange region, hand its CUDA IPC handle round and map the peers' regions."""
        region, handle = C.c_void_p(), (C.c_char * 64)()
        capi.check(self.lib.aps_k2_peer_alloc(C.byref(region), handle), "aps_k2_peer_alloc")
        handles = [None] * world
        torch.distributed.all_gather_object(handles, bytes(handle.raw))
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(region.value)
            else:
                p = C.c_void_p()
                capi.check(self.lib.aps_k2_peer_open(handles[r], C.byref(p)), "aps_k2_peer_open")
                ptrs.append(p.value)
        torch.distributed.barrier()
        return region.value, ptrs

    def peer_teardown(self, rank, own, ptrs):
        torch.cuda.synchronize()
        torch.distributed.barrier()
        for r, p in enumerate(ptrs):
            if r != rank:
                self.lib.aps_k2_peer_close(p)
        torch.distributed.barrier()
        self.lib.aps_k2_peer_free(own)

    def init(self, state, L, off, seed, density, frac_plus):
        capi.check(self.lib.aps_k2_init_device(state.data_ptr(), L, off, seed, density, frac_plus,
                                               torch.cuda.current_stream().cuda_stream), "aps_k2_init_device")

    def profile(self, state, L, off, L_global, nbins):
        cp, cm = self.zeros_i64(nbins), self.zeros_i64(nbins)
        capi.check(self.lib.aps_k2_profile_device(state.data_ptr(), L, off, L_global, nbins, cp.data_ptr(), cm.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream), "aps_k2_profile_device")
        return cp, cm

    def count(self, view):
        return int((view == 1).sum().item()), int((view == 2).sum().item())


class SublatticeLattice:
    def __init__(self, L, D, lam, beta, dt, sigma_sites=None, seed=0, backend=None, ghost=TILE, single_rank=False,
                 persistent=None):
        if L % TILE:
            raise ValueError(f"L must be a multiple of {TILE}")
        self.be = backend or CudaK2Backend()
        self.L_global, self.dt, self.seed = int(L), float(dt), int(seed)
        self.rank, self.world = (0, 1) if single_rank else dist_info()
        tiles = L // TILE
        if tiles < self.world:
            raise ValueError("fewer tiles than ranks")
        t_lo = (tiles * self.rank) // self.world
        t_hi = (tiles * (self.rank + 1)) // self.world
        self.own_lo, self.own_hi = t_lo * TILE, t_hi * TILE
        self.ghost = int(ghost) if self.world > 1 else 0
        if self.ghost % TILE:
            raise ValueError("ghost must be a multiple of the tile size")
        self.lo = max(0, self.own_lo - self.ghost)
        self.hi = min(L, self.own_hi + self.ghost)
        self.L = self.hi - self.lo
        if sigma_sites is None or sigma_sites <= 0:
            self.radius, self.w16 = -1, None           # global magnetisation: one 8-byte all-reduce per pass when sliced
        else:
            self.radius, w = fixed_point_taps(sigma_sites)
            self.w16 = self.be.from_numpy(w)
        self.rates = self.be.rates(float(D), float(lam), float(beta), float(dt))
        self.flip_tab = self.be.flip_table(self.rates) if self.radius >= 0 else None
        self.buf = [self.be.zeros_u8(self.L), self.be.zeros_u8(self.L)]
        self.cur = 0
        # persistent = one cooperative multi-pass launch per run_passes() with in-kernel exchange: always with several ranks;
        # on one GPU a launch per pass measures 5-10 % faster than a grid barrier per pass (profiles/r2_k2.md), so it is opt-in
        self.can_persist = bool(getattr(self.be, "persistent", False))
        self.persistent = self.can_persist and (self.world > 1 if persistent is None else bool(persistent))
        if self.can_persist:
            # device int64[8]: [0] grid barrier, [1] error flag, [2] cumulative own flips, [3] sum(sigma) at creation, [4] current
            self.sync = self.be.zeros_i64(8)
            self.msum = [self.sync[4:5], self.sync[5:6]]
        else:
            self.msum = [self.be.zeros_i64(1), self.be.zeros_i64(1)]
        self.n_particles = 0
        self.passes_done = 0
        # strict staleness bound: (radius + 33) sites per pass (see the module docstring); even, so that refreshes fall on step borders
        r0 = max(self.radius, 0)
        self.refresh_every = max(2, ((self.ghost - r0 - 1) // (r0 + 33)) // 2 * 2) if self.world > 1 else 1 << 30
        self.since_refresh = 0
        self._peer_own, self._peer_ptrs = None, None
        if self.persistent and self.world > 1:
            self._peer_own, self._peer_ptrs = self.be.peer_setup(self.rank, self.world)

    def close(self):
        """Unmap the peers' exchange regions and free the own one (collective: every rank must call it)."""
        if self._peer_ptrs is not None:
            self.be.peer_teardown(self.rank, self._peer_own, self._peer_ptrs)
            self._peer_own, self._peer_ptrs = None, None

    def check(self):
        """Raise if a peer time-out made the persistent kernel give up (synchronises)."""
        if self.can_persist and int(self.sync[1].item()) != 0:
            raise capi.ApsError("K2 persistent kernel: a peer did not arrive at a flag round (time-out); state is undefined")

    # ---- state ----
    @property
    def state(self):
        return self.buf[self.cur]

    def owned(self):
        a = self.own_lo - self.lo
        return self.state[a:a + (self.own_hi - self.own_lo)]

    def init_random(self, density=0.5, frac_plus=0.5):
        self.be.init(self.state, self.L, self.lo, self.seed, float(density), float(frac_plus))
        self._recount()

    def set_state(self, global_state):
        """Load a full-lattice numpy byte array (tests)."""
        self.buf[self.cur] = self.be.from_numpy(np.asarray(global_state[self.lo:self.hi], dtype=np.uint8))
        self.buf[1 - self.cur] = self.be.zeros_u8(self.L)
        self._recount()

    def set_state_from_host(self, host_state):
        """Load this rank's slab from a full-lattice uint8 tensor in (pinned) host memory: the H2D leg of the end-to-end path."""
        self.buf[self.cur].copy_(host_state[self.lo:self.hi], non_blocking=True)
        self._recount()

    def _recount(self):
        npl, nmi = self.be.count(self.owned())
        tot = torch.tensor([npl, nmi], dtype=torch.int64)
        if self.world > 1:
            t = tot.to(self.be.dev) if self.be.name == "cuda" else tot
            torch.distributed.all_reduce(t)
            tot = t.cpu()
        self.n_particles = int(tot[0] + tot[1])
        if self.can_persist:
            self.sync.zero_()
            self.sync[3] = int(tot[0] - tot[1])
            self.sync[4] = int(tot[0] - tot[1])
        elif self.radius < 0:
            self.msum[0][0] = int(tot[0] - tot[1])

    # ---- time stepping ----
    def _args(self):
        a = ApsK2Args()
        a.L, a.L_global, a.global_offset, a.n_particles = self.L, self.L_global, self.lo, max(1, self.n_particles)
        a.seed, a.pass_, a.radius = self.seed, self.passes_done, self.radius
        a.rates = self.rates
        a.w16 = self.be.ptr(self.w16) if self.w16 is not None else None
        a.flip_tab = self.be.ptr(self.flip_tab) if self.flip_tab is not None else None
        a.in_, a.out = self.be.ptr(self.buf[self.cur]), self.be.ptr(self.buf[1 - self.cur])
        a.msum_in, a.msum_out = self.be.ptr(self.msum[0]), self.be.ptr(self.msum[1])
        if self.world > 1 and self.radius < 0:          # ghost segments are recomputed by the neighbours: count own flips only
            a.count_lo, a.count_hi = self.own_lo - self.lo, self.own_hi - self.lo
        return a

    def run_passes(self, n_passes):
        if self.persistent:
            return self._run_persistent(n_passes)
        done = 0
        sliced_global = self.world > 1 and self.radius < 0
        while done < n_passes:
            k = min(n_passes - done, self.refresh_every - self.since_refresh)
            if sliced_global:
                k = 1                                  # every pass needs the lattice-wide sum(sigma) of the previous one
            a = self._args()
            self.be.run(a, k)
            if k % 2:
                self.cur = 1 - self.cur
                self.msum.reverse()
            if sliced_global:                          # msum[0] = old + own flips, msum[1] = old: all-reduce the increments
                delta = self.msum[0] - self.msum[1]
                torch.distributed.all_reduce(delta)
                self.msum[0].copy_(self.msum[1] + delta)
            self.passes_done += k
            self.since_refresh += k
            done += k
            if self.world > 1 and self.since_refresh >= self.refresh_every:
                self.exchange_ghosts()
        return self

    def _run_persistent(self, n_passes):
        """All passes in one cooperative launch; ghost refresh / sum(sigma) exchange happen inside the kernel."""
        a = self._args()
        m = capi.ApsK2Multi()
        m.n_passes, m.world, m.rank, m.refresh_every = int(n_passes), self.world, self.rank, int(min(self.refresh_every, 1 << 30))
        m.ghost, m.own_lo, m.own_hi = self.ghost, self.own_lo - self.lo, self.own_hi - self.lo
        m.buf0, m.buf1 = self.be.ptr(self.buf[self.cur]), self.be.ptr(self.buf[1 - self.cur])
        m.sync = self.be.ptr(self.sync)
        if self.world > 1:
            for r, p in enumerate(self._peer_ptrs):
                m.peer[r] = p
        self.be.run_persistent(a, m)
        if n_passes % 2:
            self.cur = 1 - self.cur
        self.passes_done += n_passes
        return self

    def run(self, n_steps):
        """n_steps full steps (two passes each): physical time advances by n_steps * dt."""
        return self.run_passes(2 * n_steps)

    def exchange_ghosts(self):
        """Refresh both ghost zones from the neighbours' owned edge sites."""
        g, st = self.ghost, self.state
        a = self.own_lo - self.lo
        own = self.own_hi - self.own_lo
        ops, recv = [], []
        P2P = torch.distributed.P2POp
        if self.rank > 0:
            send_l = st[a:a + g].clone()
            rl = torch.empty_like(send_l)
            ops += [P2P(torch.distributed.isend, send_l, self.rank - 1), P2P(torch.distributed.irecv, rl, self.rank - 1)]
            recv.append((rl, 0))
        if self.rank < self.world - 1:
            send_r = st[a + own - g:a + own].clone()
            rr = torch.empty_like(send_r)
            ops += [P2P(torch.distributed.isend, send_r, self.rank + 1), P2P(torch.distributed.irecv, rr, self.rank + 1)]
            recv.append((rr, a + own))
        for w in torch.distributed.batch_isend_irecv(ops):
            w.wait()
        for buf, at in recv:
            st[at:at + g] = buf
        self.since_refresh = 0

    # ---- observables ----
    def profile(self, nbins=1000):
        """Coarse-grained densities (particles per site) of '+' and '-' in `nbins` bins, summed over ranks."""
        a = self.own_lo - self.lo
        view = self.owned()
        cp, cm = self.be.profile(view, self.own_hi - self.own_lo, self.own_lo, self.L_global, nbins)
        if self.world > 1:
            torch.distributed.all_reduce(cp)
            torch.distributed.all_reduce(cm)
        edges = (np.arange(nbins + 1) * self.L_global + nbins - 1) // nbins     # first site of each bin (ceil)
        width = np.diff(edges).astype(float)
        return cp.cpu().numpy() / width, cm.cpu().numpy() / width

    def gather_state(self):
        """Full lattice on every rank (tests / small lattices only)."""
        own = self.owned()
        if self.world == 1:
            return own.cpu().numpy().copy()
        sizes = [((self.L_global // TILE) * (r + 1)) // self.world * TILE - ((self.L_global // TILE) * r) // self.world * TILE
                 for r in range(self.world)]
        mx = max(sizes)
        pad = torch.zeros(mx, dtype=torch.uint8, device=own.device)
        pad[: own.numel()] = own
        outs = [torch.zeros_like(pad) for _ in range(self.world)]
        torch.distributed.all_gather(outs, pad)
        return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])
