"""GraphDistribution: per-source-node categorical over out-edges, fused on the device.

Drop-in for the reference's src/reinforcement_learning.py:15-96 (`GraphDistribution(logits, edge_index,
temperature)` with `.sample()`, `.log_prob(action)`, `.entropy()`, `.mode`, `.proba`), computed by the segmented
kernels of csrc/mpnn.cu behind the C ABI, with autograd through log_prob/entropy. Declared divergences from the
literal reference code (which raises or scrambles in these cases — SURVEY.md §8c, oracle/mpnn_port.py):
D1 groups are the ranks of the distinct source ids; D3 edges inside a group are ordered by ascending edge id;
D7 a batched [B, E] input behaves row-wise like B independent 1-D distributions (sample and mode included);
the constructor's NaN assert (a host synchronisation, :19) is not performed.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.distributions import Distribution

from . import _cabi
from .topology import group_csr_for


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _action_arg(action: torch.Tensor, B: int, E: int):
    """[B, E] view of the action in a dtype the kernels read directly (no copy for bool / uint8 / int64 / fp32)."""
    a = action.reshape(B, E)
    if a.dtype == torch.bool:
        return a.view(torch.uint8), _cabi.ACTION_U8
    if a.dtype == torch.uint8:
        return a, _cabi.ACTION_U8
    if a.dtype == torch.int64:
        return a, _cabi.ACTION_I64
    return a.to(torch.float32), _cabi.ACTION_F32


def _edge_major(B: int, E: int, dtype, device, zero: bool = False):
    """A [B, E] tensor whose memory is edge-major (strides (1, B))."""
    buf = (torch.zeros if zero else torch.empty)(E, B, dtype=dtype, device=device)
    return buf.t()


class _LogProbEntropy(torch.autograd.Function):
    """(log_prob [B], entropy [B]) of logits [B,E] (any strides); `action` may be None (entropy only)."""

    @staticmethod
    def forward(ctx, logits, action, groups, temperature):
        B, E = logits.shape
        dev = logits.device
        lib = _cabi.lib()
        nb = lib.tarl_graphdist_partial_count(groups.n_rows, B)
        partials = torch.empty(max(3 * B * nb, 1), dtype=torch.float32, device=dev)
        ent = torch.empty(B, dtype=torch.float32, device=dev)
        lp = torch.empty(B, dtype=torch.float32, device=dev) if action is not None else None
        a, code = _action_arg(action, B, E) if action is not None else (None, 0)
        with torch.cuda.device(dev):
            rc = lib.tarl_graphdist_forward(groups.ref(), _cabi.rows(logits), temperature, B, _cabi.rows(a), code, None,
                                            None, ent.data_ptr(), lp.data_ptr() if lp is not None else None,
                                            partials.data_ptr(), _stream(dev))
        _cabi.check(rc, "tarl_graphdist_forward")
        ctx.groups, ctx.temperature, ctx.code = groups, temperature, code
        ctx.save_for_backward(logits, a, lp)
        if lp is None:
            lp = torch.zeros(B, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(lp)
        return lp, ent

    @staticmethod
    def backward(ctx, g_lp, g_ent):
        logits, a, lp = ctx.saved_tensors
        B, E = logits.shape
        dev = logits.device
        grad = _edge_major(B, E, torch.float32, dev) if (B > 1 and logits.stride(0) == 1) else torch.empty_like(logits)
        g_lp = g_lp.contiguous() if (g_lp is not None and a is not None) else None
        g_ent = g_ent.contiguous() if g_ent is not None else None
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_graphdist_backward(
                ctx.groups.ref(), _cabi.rows(logits), ctx.temperature, B, _cabi.rows(a), ctx.code,
                g_lp.data_ptr() if g_lp is not None else None, g_ent.data_ptr() if g_ent is not None else None,
                lp.data_ptr() if lp is not None else None, _cabi.rows(grad), _stream(dev))
        _cabi.check(rc, "tarl_graphdist_backward")
        return grad, None, None, None


class ActionSink:
    """Where a sampled action goes inside an environment: SELECTED_ROAD of R replicas, road links [R, N] and the other
    source nodes [R, N_tot - N], both contiguous and in node-id order, for the graph `edge_index` describes.
    GraphDistribution.sample(sink=...) sets `applied` when it wrote them."""

    def __init__(self, edge_index, groups, sel_links, sel_sources, n_links, n_nodes, row_offset: int = 0):
        self.edge_index, self.groups = edge_index, groups
        self.sel_links, self.sel_sources = sel_links, sel_sources
        self.n_links, self.n_nodes = int(n_links), int(n_nodes)
        self.group_node = groups.nodes.to(torch.int32).contiguous()
        self.edge_dst = edge_index[1].to(torch.int32).contiguous()
        self.applied = False
        # Which uniforms the sampling kernel draws for itself (Philox4x32-10, counter (group, global row / 4, draw)):
        self.row_offset = int(row_offset)      # global index of the first replica (ranks of a data-parallel job differ)
        self.seed_dev = None                   # int64 [1] device tensor holding the key, or None: a host seed per call
        self.draw_id = 0                       # advanced by every draw taken with the device key
        # A rollout may alternate TWO pairs of SELECTED_ROAD buffers (retarget()): prev_* are then the pair the previous
        # step's decisions live in, from which a group without a hit carries its value over.
        self.prev_links = self.prev_sources = None

    def retarget(self, sel_links, sel_sources, prev_links=None, prev_sources=None):
        self.sel_links, self.sel_sources = sel_links, sel_sources
        self.prev_links, self.prev_sources = prev_links, prev_sources

    def matches(self, dist, rows: int) -> bool:
        return (dist._groups is self.groups and self.sel_links.size(0) == rows and self.sel_links.is_contiguous()
                and (self.sel_sources is None or self.sel_sources.is_contiguous())
                and self.sel_links.device == dist._logits.device)


class GraphDistribution(Distribution):
    arg_constraints = {}
    has_rsample = False

    def __init__(self, logits: torch.Tensor, edge_index: torch.Tensor, temperature: float = 1.0):
        super().__init__(validate_args=False)
        if not logits.is_cuda:
            raise RuntimeError("GraphDistribution computes on CUDA devices only (no CPU fallback)")
        if logits.size(-1) != edge_index.size(1):
            raise ValueError("logits' last dimension must equal the number of edges")
        self.edge_index = edge_index
        self.temperature = float(temperature)
        self._lead = logits.shape[:-1]
        self._E = logits.size(-1)
        self._logits = logits.to(torch.float32).reshape(-1, self._E)       # any strides: edge-major stays edge-major
        self._groups = group_csr_for(edge_index, "source_rank")
        self.nodes = self._groups.nodes
        self.nb_nodes = self._groups.n_rows
        self._cache = {}

    # -- helpers ---------------------------------------------------------------------------------------------
    def _forward_extras(self, want_proba=False, want_mode=False):
        B, E = self._logits.shape
        dev = self._logits.device
        lg = self._logits.detach()
        em = B > 1 and lg.stride(0) == 1
        make = (lambda zero: _edge_major(B, E, torch.float32, dev, zero)) if em else \
            (lambda zero: (torch.zeros if zero else torch.empty)(B, E, dtype=torch.float32, device=dev))
        proba = make(False) if want_proba else None
        mode = make(True) if want_mode else None
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_graphdist_forward(self._groups.ref(), _cabi.rows(lg), self.temperature, B, None, 0,
                                                    _cabi.rows(proba), _cabi.rows(mode), None, None, None, _stream(dev))
        _cabi.check(rc, "tarl_graphdist_forward")
        if proba is not None:
            self._cache["proba"] = proba.reshape(*self._lead, E)
        if mode is not None:
            self._cache["mode"] = mode.reshape(*self._lead, E)

    @property
    def proba(self):
        """softmax(logits / temperature) within each source group (:25); not differentiable here — gradients flow
        through log_prob() and entropy()."""
        if "proba" not in self._cache:
            self._forward_extras(want_proba=True)
        return self._cache["proba"]

    @property
    def mode(self):
        """One-hot [.., E] of the most probable out-edge of every source node (:45-55; lowest edge id on ties)."""
        if "mode" not in self._cache:
            self._forward_extras(want_mode=True)
        return self._cache["mode"]

    @property
    def deterministic_sample(self):
        return self.mode

    # -- Distribution API ------------------------------------------------------------------------------------
    def sample(self, sample_shape=torch.Size(), uniforms: torch.Tensor | None = None, dtype=torch.int64,
               out: torch.Tensor | None = None, return_log_prob: bool = False, sink=None):
        """One-hot [sample_shape.., .., E]: inverse CDF with one uniform per (row, source group) (:57-80). int64 like
        the reference by default; dtype=torch.bool writes one byte per edge instead of eight. `uniforms` ([.., K])
        injects the noise the reference draws with torch.rand. `out` (optional): a buffer of the result's shape and
        storage dtype to write into (any strides; every entry is written), e.g. a frame of a preallocated trajectory.
        `sink` (optional, rollouts): an `ActionSink` of the environment the action is meant for. When every row shares
        one logits row and the sample is an edge-major byte one-hot, the kernel that draws the edges also writes
        SELECTED_ROAD of every replica (what SimulatorEnv._step does with the action, :223-231) and `sink.applied`
        becomes True — the environment then steps without reading the one-hot back; otherwise the sink is untouched."""
        sample_shape = torch.Size(sample_shape)
        B, E = self._logits.shape
        dev = self._logits.device
        S = int(torch.Size(sample_shape).numel()) if len(sample_shape) else 1
        lg = self._logits.detach()
        if S > 1:
            lg = lg.unsqueeze(0).expand(S, B, E).reshape(S * B, E)
        rows = S * B
        u = None                   # drawn below unless the kernel draws its own (sink path)
        if uniforms is not None:
            u = uniforms.to(device=dev, dtype=torch.float32).reshape(rows, self.nb_nodes)
        if dtype not in (torch.int64, torch.bool, torch.uint8):
            raise ValueError("sample dtype must be int64, bool or uint8")
        store = torch.int64 if dtype == torch.int64 else torch.uint8
        em = rows > 1 and lg.stride(0) in (0, 1) and len(self._lead) <= 1 and S == 1
        if out is not None:
            if out.numel() != rows * E or out.dtype not in (store, torch.bool if store == torch.uint8 else store) or out.device != dev:
                raise ValueError("out must be a buffer of the sample's shape and dtype")
            out = out.reshape(rows, E)
            out = out.view(torch.uint8) if out.dtype == torch.bool else out
        else:
            out = _edge_major(rows, E, store, dev) if em else torch.empty(rows, E, dtype=store, device=dev)
        # the log-probability of the draw comes out of the same kernel on the layout MPNNPolicyNet emits
        fused = (return_log_prob and store == torch.uint8 and rows > 1 and S == 1 and self.temperature != 0.0
                 and (rows in (4, 8, 16) or rows % 32 == 0)
                 and ((lg.stride(0) == 1 and lg.stride(1) == rows and lg.data_ptr() % 16 == 0)       # edge-major
                      or (lg.stride(0) == 0 and lg.stride(1) == 1))                                  # one row for all
                 and out.stride(0) == 1 and out.stride(1) == rows and out.data_ptr() % 4 == 0)
        lp = partials = None
        if fused:
            lp = torch.empty(rows, dtype=torch.float32, device=dev)
            partials = torch.empty(3 * rows * max(_cabi.lib().tarl_graphdist_partial_count(self.nb_nodes, rows), 1),
                                   dtype=torch.float32, device=dev)
        if sink is not None:
            sink.applied = False
        if (sink is not None and S == 1 and rows > 1 and rows % 4 == 0 and store == torch.uint8 and self.temperature != 0.0
                and lg.stride(0) == 0 and lg.stride(1) == 1 and out.stride(0) == 1 and out.stride(1) == rows
                and out.data_ptr() % 4 == 0 and sink.matches(self, rows)):
            if return_log_prob and not fused:      # any multiple of 4 rows has the fused log-probability here
                lp = torch.empty(rows, dtype=torch.float32, device=dev)
                partials = torch.empty(3 * rows * max(_cabi.lib().tarl_graphdist_partial_count(self.nb_nodes, rows), 1),
                                       dtype=torch.float32, device=dev)
                fused = True
            # no injected uniforms: the kernel draws them instead of reading back a torch.rand tensor. Key: the sink's
            # device word when it has one (rollouts replayed from a CUDA graph), else a seed from torch's default CPU
            # generator (torch.manual_seed reproduces a rollout); rows of different ranks differ through row_offset
            seed, seed_dev, draw = 0, None, 0
            if u is None:
                if sink.seed_dev is not None:
                    seed_dev, draw = sink.seed_dev.data_ptr(), sink.draw_id
                    sink.draw_id += 1
                else:
                    seed = (int(torch.randint(0, 2 ** 62, (1,))) + getattr(sink, "salt", 0)) & ((1 << 62) - 1)
            with torch.cuda.device(dev):
                rc = _cabi.lib().tarl_graphdist_sample_apply(
                    self._groups.ref(), lg.data_ptr(), self.temperature, rows, _cabi.rows(u) if u is not None else None,
                    out.data_ptr(), lp.data_ptr() if fused else None, partials.data_ptr() if fused else None,
                    sink.group_node.data_ptr(), sink.edge_dst.data_ptr(), sink.sel_links.data_ptr(),
                    sink.sel_sources.data_ptr() if sink.sel_sources is not None else None,
                    sink.prev_links.data_ptr() if sink.prev_links is not None else None,
                    sink.prev_sources.data_ptr() if sink.prev_sources is not None else None, sink.n_links, sink.n_nodes,
                    seed, seed_dev, draw, sink.row_offset, _stream(dev))
            _cabi.check(rc, "tarl_graphdist_sample_apply")
            sink.applied = True
            out = out.view(torch.bool) if dtype == torch.bool else out
            out = out.reshape(*sample_shape, *self._lead, E)
            return (out, lp.reshape(self._lead)) if return_log_prob else out
        if u is None:              # group-major memory: the rows of one group are read as one vector
            u = torch.rand(self.nb_nodes, rows, dtype=torch.float32, device=dev).t()
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_graphdist_sample(self._groups.ref(), _cabi.rows(lg), self.temperature, rows,
                                                   _cabi.rows(u), _cabi.rows(out),
                                                   _cabi.ACTION_I64 if store == torch.int64 else _cabi.ACTION_U8,
                                                   lp.data_ptr() if fused else None,
                                                   partials.data_ptr() if fused else None, _stream(dev))
        _cabi.check(rc, "tarl_graphdist_sample")
        if dtype == torch.bool:
            out = out.view(torch.bool)
        out = out.reshape(*sample_shape, *self._lead, E)
        if not return_log_prob:
            return out
        if not fused:
            return out, self.log_prob(out).detach()
        return out, lp.reshape(self._lead)

    def _lp_ent(self, action):
        return _LogProbEntropy.apply(self._logits, action, self._groups, self.temperature)

    def log_prob(self, action: torch.Tensor):
        """sum_e action_e * log(proba_e + 1e-8); -inf where a row does not select exactly one edge per group (:82-93).
        Also caches the entropy computed in the same pass for a following entropy() call."""
        if action.shape[-1] != self._E or action.numel() != self._logits.numel():
            raise ValueError("action must have the shape of logits")
        lp, ent = self._lp_ent(action.to(self._logits.device))
        self._cache["entropy"] = ent
        return lp.reshape(self._lead)

    def entropy(self):
        """-sum_e proba_e * log(proba_e + 1e-8), flattened to [B] (:95-96)."""
        ent = self._cache.pop("entropy", None)
        if ent is None:
            _, ent = self._lp_ent(None)
        return ent.flatten()
