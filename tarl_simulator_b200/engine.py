"""LinkStore: the resident, compact device representation of the road state (csrc/engine.cu).

It steps R independent replicas of one network with two kernel launches per timestep and converts from/to the
reference's `graph.x[:num_roads]` row layout exactly (`from_graph` / `export_x`). It is what loops that keep the
state on the device between steps use (benchmarks, batched PPO rollouts); the per-call drop-in
`SimulationCoreModel.forward(graph)` works on `graph.x` directly (csrc/core_step.cu).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from .topology import topology_for

PHASE_SELECT_APPEND, PHASE_RESPOND_POP = 2, 4
VARIANT_ELL, VARIANT_CSR = 0, 1    # kernel families of tarl_store_step (bit-identical results)
CLUSTER = 128                      # links per locality cluster = threads per CTA of the step kernels


def locality_order(edge_index_routes: torch.Tensor, n_links: int, cluster: int = CLUSTER) -> torch.Tensor:
    """slot -> link id: clusters of `cluster` links grown breadth-first over the dual graph (tarl_cluster_links, host
    C++), so that the neighbour gathers of a CTA tile mostly hit lines the tile itself loads. One-off preprocessing."""
    ei = edge_index_routes.detach().cpu().numpy().astype(np.int64)
    a = np.concatenate([ei[0], ei[1]])
    b = np.concatenate([ei[1], ei[0]])
    order = np.argsort(a, kind="stable")
    idx = np.ascontiguousarray(b[order].astype(np.int32))
    ptr = np.zeros(n_links + 1, dtype=np.int32)
    np.cumsum(np.bincount(a, minlength=n_links), out=ptr[1:])
    out = np.empty(max(n_links, 1), dtype=np.int32)
    rc = _cabi.lib().tarl_cluster_links(n_links, ptr.ctypes.data, idx.ctypes.data if idx.size else None, cluster,
                                        out.ctypes.data)
    _cabi.check(rc, "tarl_cluster_links")
    return torch.from_numpy(out[:n_links])


class LinkStore:
    def __init__(self, edge_index_routes: torch.Tensor, edge_attr_routes: torch.Tensor, n_links: int, Nmax: int,
                 replicas: int = 1, device=None, seed: int = 0, cluster: bool | None = None):
        """cluster: keep the links in a locality order inside the store (tarl_cluster_links). Purely internal: every
        input and output of this class stays in link-id order. Off by default: measured on B200 it raises the share
        of neighbour gathers served inside a CTA tile from 48 % to 71 % on the 1M-link ring-radial network but costs
        more (delta_tt no longer lands in source order) than the L1 hits return (69.8 vs 61.7 us per step)."""
        dev = torch.device(device) if device is not None else edge_index_routes.device
        if dev.type != "cuda":
            raise RuntimeError("LinkStore lives on a CUDA device (no CPU fallback)")
        self.device, self.N, self.R, self.Nmax, self.M = dev, int(n_links), int(replicas), int(Nmax), int(Nmax) - 1
        if cluster is None:
            cluster = False
        self.slot_link = self.link_slot = None
        if cluster and self.N > 0 and edge_index_routes.size(1) > 0:
            order = locality_order(edge_index_routes, self.N)
            self.slot_link = order.to(dev, torch.int32)
            inv = torch.empty(self.N, dtype=torch.int64)
            inv[order.long()] = torch.arange(self.N)
            self.link_slot = inv.to(dev, torch.int32)
            self._ei_slots = self.link_slot.long()[edge_index_routes.to(dev).long()]    # endpoints renamed, edge ids kept
            edge_index_routes = self._ei_slots
        self.topo = topology_for(edge_index_routes, self.N)
        self.E = self.topo.n_edges
        attr = edge_attr_routes.reshape(-1).to(torch.float32)
        self.attr_in = attr[self.topo.in_eid.long()].contiguous()        # edge_attr in CSR-by-target order
        self._ell = self.topo.ell(edge_attr_routes)                      # ELLPACK copy of the first W edges per link
        # stat_a.w = the weight all in-edges of a link share (NaN where they differ), re-applied after every import:
        # the direction kernel then skips the edge-weight column of those links. TARL_NO_UNIFORM_WEIGHTS=1 (tests,
        # A/B measurements) leaves the hint off: same results, every link reads its column.
        self.uniform_weights = not os.environ.get("TARL_NO_UNIFORM_WEIGHTS")
        L = self.N * self.R
        f32 = dict(dtype=torch.float32, device=dev)
        self.hot = [torch.zeros(max(L, 1), 8, **f32), torch.zeros(max(L, 1), 8, **f32)]
        self.cur = 0
        self.sel = torch.zeros(max(L, 1), **f32)
        self.stat_a = torch.zeros(max(self.N, 1), 4, **f32)
        self.stat_b = torch.zeros(max(self.N, 1), 4, **f32)
        self.queue = torch.zeros(max(L * self.M, 1), 4, **f32)
        self.post = torch.zeros(max(L, 1), 2, **f32)
        self.hint = torch.zeros(max(L, 1), dtype=torch.uint8, device=dev)
        self.pop = torch.zeros(max(L, 1), dtype=torch.uint8, device=dev)
        self.words = (self.N + 31) // 32                              # pop mask words per replica (bit form)
        self.pop_bits = torch.zeros(max(self.R * self.words, 1), dtype=torch.int32, device=dev)
        self.dtt_link = torch.zeros(max(L, 1), **f32)                 # delta_travel_time per (replica, upstream link)
        self.flags = torch.zeros(_cabi.FLAG_COUNT, dtype=torch.int32, device=dev)
        self._io = _cabi.StepIO()
        self.seed, self.step_id = int(seed), 0
        self.seed_dev = None          # optional int64 [1] device tensor: the key of the in-kernel noise lives there
        self.t_last = 0.0
        self._struct = _cabi.LinkStore()
        self._fill_struct()

    def _fill_struct(self):
        s = self._struct
        s.n_links, s.n_replicas, s.nmax = self.N, self.R, self.Nmax
        s.hints = _cabi.STORE_UNIFORM_WEIGHTS if self.uniform_weights else 0
        s.hot_cur, s.hot_next = self.hot[self.cur].data_ptr(), self.hot[self.cur ^ 1].data_ptr()
        s.sel, s.stat_a, s.stat_b = self.sel.data_ptr(), self.stat_a.data_ptr(), self.stat_b.data_ptr()
        s.queue, s.post, s.pop_hint = self.queue.data_ptr(), self.post.data_ptr(), self.hint.data_ptr()
        s.slot_link = self.slot_link.data_ptr() if self.slot_link is not None else None
        s.link_slot = self.link_slot.data_ptr() if self.link_slot is not None else None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_graph(cls, graph, Nmax: int, replicas: int = 1, seed: int = 0, cluster: bool | None = None) -> "LinkStore":
        """Build from a reference-layout graph (graph.x on a CUDA device); every replica starts from graph.x."""
        N = int(graph.num_roads)
        store = cls(graph.edge_index_routes, graph.edge_attr_routes, N, Nmax, replicas, graph.x.device, seed, cluster)
        cc = graph.congestion_constant[:N] if hasattr(graph, "congestion_constant") and hasattr(graph, "critical_number") else None
        store.import_x(graph.x[:N], cc, broadcast=True)
        return store

    def import_x(self, x: torch.Tensor, cc: torch.Tensor | None = None, broadcast: bool = False):
        """x: [N, F] (broadcast=True: the same rows for every replica) or [R, N, F]; fp32 rows with stride(-1) == 1."""
        F = 3 * self.Nmax + 7
        if x.dtype != torch.float32 or x.size(-1) != F or (x.numel() and x.stride(-1) != 1) or x.device != self.device:
            raise ValueError(f"x must be fp32 [.., N, {F}] with contiguous rows on {self.device}")
        if x.dim() == 2:
            if not (broadcast or self.R == 1):
                raise ValueError("2-D x needs broadcast=True when replicas > 1")
            row_stride, rep_stride = x.stride(0) if self.N > 1 else F, 0
        else:
            if x.size(0) != self.R:
                raise ValueError("leading dimension of x must equal the number of replicas")
            row_stride, rep_stride = x.stride(1) if self.N > 1 else F, x.stride(0) if self.R > 1 else 0
        ccp = None
        if cc is not None:
            self._cc = cc.to(torch.float32).contiguous()
            ccp = self._cc.data_ptr()
        self._fill_struct()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_import(C.byref(self._struct), x.data_ptr(), row_stride, rep_stride, ccp,
                                               self.flags.data_ptr(), self._stream())
        _cabi.check(rc, "tarl_store_import")
        if self.uniform_weights and self.N > 0:
            self.stat_a[: self.N, 3] = self._ell[4][: self.N]

    def export_x(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """The road rows exactly as the reference would hold them: [R, N, F] (or into `out`, [N,F] allowed if R==1)."""
        F = 3 * self.Nmax + 7
        if out is None:
            out = torch.empty(self.R, self.N, F, dtype=torch.float32, device=self.device)
        x3 = out if out.dim() == 3 else out.unsqueeze(0)
        if x3.size(0) != self.R or x3.size(1) != self.N or x3.size(2) != F or x3.dtype != torch.float32 or (x3.numel() and x3.stride(2) != 1):
            raise ValueError("bad export target")
        self._fill_struct()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_export(C.byref(self._struct), x3.data_ptr(), x3.stride(1) if self.N > 1 else F,
                                               x3.stride(0) if self.R > 1 else 0, float(self.t_last), self._stream())
        _cabi.check(rc, "tarl_store_export")
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _to_slots(self, per_link: torch.Tensor) -> torch.Tensor:
        """[.., N] in link-id order -> the store's slot order."""
        return per_link if self.slot_link is None else per_link[..., self.slot_link.long()]

    def _to_links(self, per_slot: torch.Tensor) -> torch.Tensor:
        return per_slot if self.link_slot is None else per_slot[..., self.link_slot.long()]

    def set_selected_road(self, sel: torch.Tensor):
        """This step's SELECTED_ROAD values, [N] (all replicas) or [R, N], in link-id order."""
        self.sel[: self.R * self.N].view(self.R, self.N).copy_(self._to_slots(sel.to(torch.float32).reshape(-1, self.N)))

    def selected_road(self) -> torch.Tensor:
        """SELECTED_ROAD [R, N] in link-id order."""
        return self._to_links(self.sel[: self.R * self.N].view(self.R, self.N))

    def _step_io(self, t, noise, step_id, want_dtt: bool, want_bits: bool, out=None):
        """out (optional): dict of caller-owned per-step output buffers replacing the store's own — "pop" (uint8 / bool
        [R*N]), "flags" (int32 [FLAG_COUNT], zeroed), "delta_tt_link" (fp32 [R*N]), "pop_bits" (int32 [R*words])."""
        io = self._io
        out = out or {}
        io.noise = noise.data_ptr() if noise is not None else None
        io.seed, io.step_id, io.t = self.seed, step_id, float(t)
        io.seed_dev = self.seed_dev.data_ptr() if self.seed_dev is not None else None
        dtt = out.get("delta_tt_link")
        io.delta_tt_link = dtt.data_ptr() if dtt is not None else (self.dtt_link.data_ptr() if want_dtt else None)
        bits = out.get("pop_bits")
        io.pop = out["pop"].data_ptr() if out.get("pop") is not None else self.pop.data_ptr()
        io.pop_bits = bits.data_ptr() if bits is not None else (self.pop_bits.data_ptr() if want_bits else None)
        io.flags = out["flags"].data_ptr() if out.get("flags") is not None else self.flags.data_ptr()
        return C.byref(io)

    def expand_delta_tt(self, out: torch.Tensor | None = None, per_link: torch.Tensor | None = None) -> torch.Tensor:
        """road_optimality_data["delta_travel_time"] of the latest step in the reference's form: [R, E] fp32 in
        original edge order, out[r, e] = delta_tt_link[r, source link of e] (tarl_expand_delta_tt). per_link: the
        [R*N] per-link values to expand (default: the store's own buffer, i.e. the latest step's)."""
        if out is None:
            out = torch.empty(self.R, self.E, dtype=torch.float32, device=self.device)
        if out.numel() != self.R * self.E or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise ValueError("delta_tt must be a contiguous fp32 [R, E] tensor")
        src = self.dtt_link if per_link is None else per_link
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_expand_delta_tt(self.topo.src32.data_ptr(), self.E, self.N, self.R,
                                                  src.data_ptr(), out.data_ptr(), self._stream())
        _cabi.check(rc, "tarl_expand_delta_tt")
        return out

    def delta_tt_link(self) -> torch.Tensor:
        """delta_travel_time of the latest step per (replica, upstream link), link-id order: [R, N]."""
        return self._to_links(self.dtt_link[: self.R * self.N].view(self.R, self.N))

    def noise_of_step(self, step_id: int | None = None) -> torch.Tensor:
        """The [R, E] uniforms (original edge order) the in-kernel Philox stream yields for `step_id` (default: the
        next step): step(noise=that) is bit-identical to step(noise=None), and the CPU oracle can replay it."""
        out = torch.full((self.R, self.E), 0.5, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_noise(self.topo.ref(), self.R, self.seed,
                                              self.step_id if step_id is None else int(step_id), out.data_ptr(),
                                              self._stream())
        _cabi.check(rc, "tarl_store_noise")
        return out

    def step(self, t: float, noise: torch.Tensor | None = None, delta_tt: torch.Tensor | None = None,
             phase_mask: int = PHASE_SELECT_APPEND | PHASE_RESPOND_POP, variant: int = VARIANT_ELL, withdraw=None,
             delta_tt_link: bool = False, pop_bits: bool = False, out=None):
        """One core step for all replicas. noise: [R, E] (or [E] when R == 1) uniforms in original edge order, or None
        for the in-kernel Philox stream. delta_tt: optional [R, E] output in the reference's per-edge form (one more
        launch: the step itself emits one value per upstream link, self.dtt_link — delta_tt_link=True asks for that
        alone). pop_bits=True also leaves the pop mask as bits in self.pop_bits. Returns the pop mask view [R, N]
        (uint8).
        withdraw (optional, ELL variant and link-id order only — see can_fuse_withdraw()): dict(table=_cabi.AgentTable,
        adjacency=CSR struct, n_nodes, mask, counters, num_out, occupancy) — the withdrawal at the same t and the
        occupancy observation ride on the response phase (tarl_store_step_withdraw)."""
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
            if noise.numel() != self.R * self.E:
                raise ValueError("noise must hold one uniform per replica and dual edge")
        if delta_tt is not None and (delta_tt.numel() != self.R * self.E or delta_tt.dtype != torch.float32 or not delta_tt.is_contiguous()):
            raise ValueError("delta_tt must be a contiguous fp32 [R, E] tensor")
        self._fill_struct()
        io = self._step_io(t, noise, self.step_id, delta_tt is not None or delta_tt_link, pop_bits, out)
        if withdraw is not None:
            if not self.can_fuse_withdraw(variant, phase_mask):
                raise ValueError("withdraw= needs the ELL variant, both phases and a store in link-id order")
            w = withdraw
            with torch.cuda.device(self.device):
                rc = _cabi.lib().tarl_store_step_withdraw(
                    self.topo.ref(), C.byref(self._ell[0]), C.byref(self._struct), self.attr_in.data_ptr(), io,
                    C.byref(w["table"]), C.byref(w["adjacency"]), int(w["n_nodes"]), w["mask"].data_ptr(),
                    w["counters"].data_ptr() if w.get("counters") is not None else None, w["num_out"].data_ptr(),
                    w["occupancy"].data_ptr(), self._stream())
            _cabi.check(rc, "tarl_store_step_withdraw")
        else:
            with torch.cuda.device(self.device):
                rc = _cabi.lib().tarl_store_step(
                    self.topo.ref(), C.byref(self._ell[0]) if variant == VARIANT_ELL else None, C.byref(self._struct),
                    self.attr_in.data_ptr(), io, self._stream(), phase_mask)
            _cabi.check(rc, "tarl_store_step")
        if delta_tt is not None and (phase_mask & PHASE_SELECT_APPEND):
            self.expand_delta_tt(delta_tt)
        if phase_mask & PHASE_RESPOND_POP:
            self.cur ^= 1
            self.step_id += 1
            self.t_last = float(t)
        return self._to_links(self.pop[: self.N * self.R].view(self.R, self.N))

    # ------------------------------------------------------------------------------------------------------------
    def _host_pipe(self):
        """The pipe of tarl_store_step_host (two copy streams + events, owned by the library) and the per-slot device
        buffers a host step uses: SELECTED_ROAD staging, delta_travel_time per link, pop bits."""
        if getattr(self, "_pipe", None) is None:
            if self.slot_link is not None:
                raise NotImplementedError("host steps need the store in link-id order (cluster=False)")
            h = C.c_void_p()
            _cabi.check(_cabi.lib().tarl_host_pipe_create(C.byref(h)), "tarl_host_pipe_create")
            self._pipe = h
            L = max(self.N * self.R, 1)
            f32 = dict(dtype=torch.float32, device=self.device)
            self._host_stage = [torch.empty(L, **f32) for _ in range(2)]
            self._host_dtt = [torch.empty(L, **f32) for _ in range(2)]
            self._host_bits = [torch.empty(max(self.R * self.words, 1), dtype=torch.int32, device=self.device) for _ in range(2)]
            self._host_slot = 0
        return self._pipe

    def step_host(self, t: float, sel_host: torch.Tensor | None = None, dtt_host: torch.Tensor | None = None,
                  pop_bits_host: torch.Tensor | None = None, noise: torch.Tensor | None = None, out=None,
                  variant: int = VARIANT_ELL):
        """One core step with HOST inputs / outputs (tarl_store_step_host): sel_host (pinned fp32 [R*N], this step's
        SELECTED_ROAD in link-id order), dtt_host (pinned fp32 [R*N]) and pop_bits_host (pinned int32 [R*words]) are
        copied on the pipe's own streams around the kernels; successive calls alternate two slots, so the upload of the
        next step and the download of the previous one overlap this step's kernels. Returns the dict of this step's
        DEVICE outputs ("delta_tt_link", "pop_bits" — per-slot buffers, overwritten by the next-but-one host step —
        plus whatever `out` held). The host buffers are complete once host_join() + a synchronisation of the current
        stream (or of the device) has passed."""
        pipe = self._host_pipe()
        slot = self._host_slot
        self._host_slot ^= 1
        o = dict(out) if out else {}
        o.setdefault("delta_tt_link", self._host_dtt[slot])
        o.setdefault("pop_bits", self._host_bits[slot])
        stage = None
        if sel_host is not None:
            stage = self._host_stage[slot]
            self.sel = stage                                  # this step's SELECTED_ROAD array (a pointer swap)
        self._fill_struct()
        io = self._step_io(t, noise, self.step_id, True, True, o)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_step_host(
                self.topo.ref(), C.byref(self._ell[0]) if variant == VARIANT_ELL else None, C.byref(self._struct),
                self.attr_in.data_ptr(), io, pipe, slot, sel_host.data_ptr() if sel_host is not None else None,
                stage.data_ptr() if stage is not None else None, dtt_host.data_ptr() if dtt_host is not None else None,
                pop_bits_host.data_ptr() if pop_bits_host is not None else None, self._stream())
        _cabi.check(rc, "tarl_store_step_host")
        self.cur ^= 1
        self.step_id += 1
        self.t_last = float(t)
        return o

    def host_join(self):
        """torch's current stream waits for every host copy the host steps have issued."""
        if getattr(self, "_pipe", None) is not None:
            with torch.cuda.device(self.device):
                _cabi.check(_cabi.lib().tarl_host_pipe_join(self._pipe, self._stream()), "tarl_host_pipe_join")

    def __del__(self):
        pipe = getattr(self, "_pipe", None)
        if pipe is not None:
            try:
                _cabi.lib().tarl_host_pipe_destroy(pipe)
            except Exception:
                pass
            self._pipe = None

    def can_fuse_withdraw(self, variant: int = VARIANT_ELL, phase_mask: int = PHASE_SELECT_APPEND | PHASE_RESPOND_POP) -> bool:
        return (variant == VARIANT_ELL and self.slot_link is None and self.N > 0
                and phase_mask == (PHASE_SELECT_APPEND | PHASE_RESPOND_POP))

    def run(self, t0: float, n_steps: int, dt: float = 1.0, sel_bank=None, delta_tt: torch.Tensor | None = None,
            variant: int = VARIANT_ELL, delta_tt_link: bool = True, pop_bits: bool = False):
        """`n_steps` consecutive core steps enqueued by ONE call into the library (tarl_store_run; in-kernel noise).
        sel_bank: optional list of fp32 [R*N] device tensors cycled through as successive steps' SELECTED_ROAD.
        Every step writes delta_travel_time per upstream link (self.dtt_link) and the pop mask; delta_tt ([R, E]) is
        materialised once, from the last step."""
        ptrs = None
        nb = 0
        if sel_bank:
            nb = len(sel_bank)
            for b in sel_bank:
                if b.dtype != torch.float32 or b.numel() != self.R * self.N or not b.is_contiguous() or b.device != self.device:
                    raise ValueError("sel_bank entries must be contiguous fp32 [R*N] tensors on the store's device")
            if self.slot_link is not None:      # link-id order -> slot order, once per bank tensor
                cache = self.__dict__.setdefault("_bank_cache", {})
                for b in sel_bank:
                    key = (b.data_ptr(), b._version)
                    if key not in cache:
                        if len(cache) > 64:
                            cache.clear()
                        cache[key] = self._to_slots(b.view(self.R, self.N)).contiguous().view(-1)
                sel_bank = [cache[(b.data_ptr(), b._version)] for b in sel_bank]
            ptrs = (C.c_void_p * nb)(*[b.data_ptr() for b in sel_bank])
        self._fill_struct()
        io = self._step_io(t0, None, self.step_id, delta_tt is not None or delta_tt_link, pop_bits)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_store_run(
                self.topo.ref(), C.byref(self._ell[0]) if variant == VARIANT_ELL else None, C.byref(self._struct),
                self.attr_in.data_ptr(), io, float(dt), int(n_steps), ptrs, nb, self._stream())
        _cabi.check(rc, "tarl_store_run")
        if delta_tt is not None and n_steps > 0:
            self.expand_delta_tt(delta_tt)          # the LAST step's values
        if n_steps > 0:
            self.cur ^= n_steps & 1
            self.step_id += n_steps
            self.t_last = float(t0) + float(dt) * (n_steps - 1)
            if sel_bank:
                self.sel = sel_bank[(n_steps - 1) % nb]
        return self._to_links(self.pop[: self.N * self.R].view(self.R, self.N))

    def clear_queues(self):
        """TransportationSimulator.reset (src/transportation_simulator.py:353-358) on the store: the three queue
        segments and NUM become zero on every link; MAXN, the statics and SELECTED_ROAD stay."""
        hot = self.hot[self.cur]
        hot[:, 0:3] = 0.0
        hot[:, 4:8] = 0.0          # head arrival, tail id, pending garbage, meta (ring head 0, no garbage)
        self.queue.zero_()

    def num_agents(self) -> torch.Tensor:
        """NUMBER_OF_AGENT per (replica, link) in link-id order (a strided view into the hot records when the store
        keeps the links in their own order)."""
        return self._to_links(self.hot[self.cur][: self.N * self.R, 2].view(self.R, self.N))

    def head_agents(self) -> torch.Tensor:
        """Head agent id per (replica, link), fp32 as stored, link-id order."""
        return self._to_links(self.hot[self.cur][: self.N * self.R, 0].view(self.R, self.N))

    def check_errors(self):
        bits = int(self.flags[_cabi.FLAG_ERROR])
        if bits:
            raise RuntimeError("link store fault: " + _cabi.decode_error_bits(bits))
