"""The per-timestep network step behind the reference's own classes.

Drop-in for src/simulation_core_model.py, src/direction_mpnn.py and src/response_mpnn.py of the reference: same class
names, constructor arguments, `forward` signatures, side outputs (`road_optimality_data["delta_travel_time"]`,
`update_history`) and in-place mutation of `graph.x[:num_roads]`. The arithmetic is three CUDA kernels behind the C
ABI (csrc/core_step.cu); there is no PyTorch or CPU implementation of it in this package.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _cabi
from ._arena import StepArena as _StepArena
from .feature_helpers import FeatureHelpers
from .message_passing import MessagePassing
from .topology import topology_for


def _require_cuda_rows(x: torch.Tensor, Nmax: int):
    if not x.is_cuda:
        raise RuntimeError("tarl_simulator_b200 computes on CUDA devices only (no CPU fallback): move graph.x to cuda")
    if x.dtype != torch.float32 or x.dim() != 2 or x.size(1) != 3 * Nmax + 7:
        raise ValueError(f"x must be fp32 [N, 3*Nmax+7={3 * Nmax + 7}], got {x.dtype} {tuple(x.shape)}")
    if x.size(0) > 1 and x.stride(1) != 1:
        raise ValueError("x rows must be contiguous (stride(1) == 1)")


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Scratch:
    """Per-module device scratch, grown on demand (caller-owned memory of the C ABI)."""

    def __init__(self):
        self.buf = None

    def get(self, n_links: int, device) -> torch.Tensor:
        need = _cabi.lib().tarl_core_workspace_bytes(n_links)
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        return self.buf


class UpdateHistory(list):
    """`ResponseMPNN.update_history`: list of (time, bool[N]) for the steps on which at least one link popped
    (src/response_mpnn.py:106-107,125). Whether a step popped anything is known on the device only; entries are
    therefore queued with their device flag and admitted to the list when the list is next read (or every
    `flush_every` steps), so that stepping never forces a host synchronisation. Kernel error bits ride along and
    raise here."""

    flush_every = 64

    def __init__(self, items=()):
        super().__init__(items)
        self._pending = []
        self.keep = True        # False: masks are not retained (on-device counters instead); error bits still raise

    def push(self, time, mask: torch.Tensor, flags: torch.Tensor):
        self._pending.append((time, mask, flags))
        if len(self._pending) >= self.flush_every:
            self.resolve()

    def resolve(self):
        if not self._pending:
            return
        pend, self._pending = self._pending, []
        host = torch.stack([f for _, _, f in pend]).cpu()
        err = 0
        for (time, mask, _), fl in zip(pend, host.tolist()):
            err |= fl[_cabi.FLAG_ERROR]
            if fl[_cabi.FLAG_ANY_POP] and self.keep:
                super().append((time, mask))
        if err:
            raise RuntimeError("core step fault: " + _cabi.decode_error_bits(err))

    def _r(self):
        self.resolve()
        return self

    def __len__(self):
        self.resolve(); return super().__len__()

    def __iter__(self):
        self.resolve(); return super().__iter__()

    def __getitem__(self, i):
        self.resolve(); return super().__getitem__(i)

    def __bool__(self):
        return len(self) > 0

    def __add__(self, other):
        self.resolve(); return list(super().__iter__()) + list(other)

    def __radd__(self, other):
        self.resolve(); return list(other) + list(super().__iter__())

    def __eq__(self, other):
        self.resolve(); return list(super().__iter__()) == list(other)

    def __repr__(self):
        self.resolve(); return super().__repr__()


class DirectionMPNN(MessagePassing, FeatureHelpers):
    """Admission + hand-off (src/direction_mpnn.py:19-236): per dual edge eligibility, per downstream link a
    Gumbel-max pick of ONE upstream head ∝ edge_attr, tail append with congestion-dependent exit time on every link.

    `forward(x, edge_index, edge_attr, critical_number=None, congestion_constant=None, noise=None)` mutates `x` in
    place and returns it. `noise` (optional, [E] uniforms in original edge order) replaces the draw the reference
    makes with torch.rand_like (src/direction_mpnn.py:137); when omitted, E uniforms are drawn from torch's global
    generator on x.device — the same stream the reference would consume on that device."""

    def __init__(self, Nmax=100, time: int = 0):
        MessagePassing.__init__(self)
        FeatureHelpers.__init__(self, Nmax=Nmax)
        self.time = time
        self.Nmax = Nmax
        self.road_optimality_data = None
        self._scratch = _Scratch()
        self._flags = None

    def set_time(self, time):
        self.time = time

    def _flag_words(self, device):
        return torch.zeros(_cabi.FLAG_COUNT, dtype=torch.int32, device=device)

    def forward(self, x, edge_index, edge_attr, critical_number: Optional[torch.Tensor] = None,
                congestion_constant: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        _require_cuda_rows(x, self.Nmax)
        N, E = x.size(0), edge_index.size(1)
        topo = topology_for(edge_index, N)
        attr, cc, noise = _edge_inputs(x, E, edge_attr, critical_number, congestion_constant, noise)
        delta_tt = torch.empty(E, dtype=torch.float32, device=x.device)
        flags = self._flag_words(x.device)
        ws = self._scratch.get(N, x.device)
        with torch.cuda.device(x.device):
            rc = _cabi.lib().tarl_direction_forward(
                topo.ref(), x.data_ptr(), x.stride(0) if N > 1 else x.size(1), self.Nmax, attr.data_ptr(),
                cc.data_ptr() if cc is not None else None, noise.data_ptr(), None, float(self.time),
                delta_tt.data_ptr(), flags.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x.device))
        _cabi.check(rc, "tarl_direction_forward")
        self.road_optimality_data = {"delta_travel_time": delta_tt}
        self._flags = flags
        return x

    def check_errors(self):
        """Synchronises and raises if the last forward hit a data-dependent fault."""
        if self._flags is not None:
            bits = int(self._flags[_cabi.FLAG_ERROR])
            if bits:
                raise RuntimeError("direction step fault: " + _cabi.decode_error_bits(bits))


class ResponseMPNN(MessagePassing, FeatureHelpers):
    """Acknowledgement (src/response_mpnn.py:6-139): an upstream link pops its FIFO head iff the tail of one of its
    downstream links now equals that head; the three queue segments shift left by one, NUM -= 1, and the pop mask
    joins `update_history` on steps where anything popped."""

    def __init__(self, Nmax: int = 100, time: int = 0):
        MessagePassing.__init__(self, aggr="max", flow="target_to_source")
        FeatureHelpers.__init__(self, Nmax=Nmax)
        self.time = time
        self._history = UpdateHistory()
        self._scratch = _Scratch()

    @property
    def update_history(self):
        return self._history

    @update_history.setter
    def update_history(self, value):
        self._history = value if isinstance(value, UpdateHistory) else UpdateHistory(value)

    def set_time(self, time: int):
        self.time = time

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor = None) -> torch.Tensor:
        _require_cuda_rows(x, self.Nmax)
        N = x.size(0)
        topo = topology_for(edge_index, N)
        pop = torch.empty(N, dtype=torch.bool, device=x.device)
        flags = torch.zeros(_cabi.FLAG_COUNT, dtype=torch.int32, device=x.device)
        ws = self._scratch.get(N, x.device)
        with torch.cuda.device(x.device):
            rc = _cabi.lib().tarl_response_forward(
                topo.ref(), x.data_ptr(), x.stride(0) if N > 1 else x.size(1), self.Nmax, pop.data_ptr(),
                flags.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x.device))
        _cabi.check(rc, "tarl_response_forward")
        self._history.push(self.time, pop, flags)
        return x


def _edge_inputs(x, E, edge_attr, critical_number, congestion_constant, noise):
    attr = edge_attr.reshape(-1)
    if attr.numel() != E or attr.dtype != torch.float32 or attr.device != x.device:
        raise ValueError("edge_attr must be fp32 [E,1] on x's device")
    attr = attr.contiguous()
    cc = None
    if critical_number is not None and congestion_constant is not None:   # src/direction_mpnn.py:179
        cc = congestion_constant.to(torch.float32).contiguous()
        if cc.numel() != x.size(0) or cc.device != x.device:
            raise ValueError("congestion_constant must have one entry per row of x, on x's device")
    if noise is None:
        noise = torch.rand(E, dtype=torch.float32, device=x.device)
    else:
        noise = noise.to(device=x.device, dtype=torch.float32).contiguous()
        if noise.numel() != E:
            raise ValueError("noise must hold one uniform per dual edge")
    return attr, cc, noise


def _sel_ptr(sel, N, dev):
    if sel is None:
        return None
    if sel.dtype != torch.float32 or sel.device != dev or sel.numel() != N or not sel.is_contiguous():
        raise ValueError("selected_road must be a contiguous fp32 [N] tensor on x's device")
    return sel.data_ptr()


class SimulationCoreModel(nn.Module):
    """One network timestep on the road sub-graph (src/simulation_core_model.py:10-88): DirectionMPNN then
    ResponseMPNN, in place on `graph.x[:graph.num_roads]`. Insertion and withdrawal of agents are not part of it.

    Parameters mirror the reference: `Nmax`, `device`, `time`, `torch_compile` (accepted and ignored: there is no
    tracing compiler on this path). `forward(graph, noise=None)` additionally accepts the E uniforms to inject."""

    def __init__(self, Nmax: int, device: str, time: int, torch_compile: bool = False):
        super().__init__()
        self.direction_mpnn = DirectionMPNN(Nmax=Nmax, time=time)
        self.response_mpnn = ResponseMPNN(Nmax=Nmax, time=time)
        self.time = time
        self.Nmax = Nmax
        self.device = device
        self.last_pop = None            # bool[N] of the latest step (device), whether or not it joined the history
        self._scratch = _Scratch()
        self._arena = _StepArena()

    def set_time(self, time):
        self.time = time
        self.direction_mpnn.set_time(time)
        self.response_mpnn.set_time(time)

    def forward(self, graph, noise: Optional[torch.Tensor] = None, selected_road: Optional[torch.Tensor] = None):
        """`selected_road` (optional, fp32 [N] on the device): this step's SELECTED_ROAD column, applied inside the
        first kernel instead of by a separate strided write into graph.x beforehand."""
        N = int(graph.num_roads)
        x_roads = graph.x[:N]                       # a view: every write lands in graph.x (reference :52,:81)
        _require_cuda_rows(x_roads, self.Nmax)
        ei = graph.edge_index_routes
        E = ei.size(1)
        topo = topology_for(ei, N)
        has_static = hasattr(graph, "critical_number") and hasattr(graph, "congestion_constant")
        attr, cc, noise = _edge_inputs(
            x_roads, E, graph.edge_attr_routes,
            graph.critical_number[:N] if has_static else None,
            graph.congestion_constant[:N] if has_static else None, noise)
        dev = x_roads.device
        delta_tt = torch.empty(E, dtype=torch.float32, device=dev)
        pop, flags = self._arena.take(N, dev)
        ws = self._scratch.get(N, dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_core_step(
                topo.ref(), x_roads.data_ptr(), x_roads.stride(0) if N > 1 else x_roads.size(1), self.Nmax,
                attr.data_ptr(), cc.data_ptr() if cc is not None else None, noise.data_ptr(),
                _sel_ptr(selected_road, N, dev), float(self.time),
                delta_tt.data_ptr(), pop.data_ptr(), flags.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev))
        _cabi.check(rc, "tarl_core_step")
        self.direction_mpnn.road_optimality_data = {"delta_travel_time": delta_tt}
        self.direction_mpnn._flags = flags
        self.response_mpnn.update_history.push(self.time, pop, flags)
        self.last_pop = pop
        return graph

    def check_errors(self):
        """Synchronises; raises if any queued step reported a data-dependent fault."""
        self.response_mpnn.update_history.resolve()
