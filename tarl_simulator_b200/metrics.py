"""LinkMetrics: the simulator's metrics side channels, accumulated on the device (csrc/metrics.cu).

The reference keeps one bool[N] per step in `ResponseMPNN.update_history` and `Agents.withdraw_history` and one host
copy of `delta_travel_time[E]` per step in `road_optimality_values`, and reduces them afterwards in
`TransportationSimulator.compute_node_metrics` / `plot_daily_counts` / `plot_road_optimality`
(src/transportation_simulator.py:351,453-510,563-669). At a million links that is 1 MB + 16 MB per simulated second.
Here the same reductions run as the masks are produced: hourly integer counters [R, H, N] and, optionally, hourly
sums and the latest value of the per-link road-optimality aggregate. `node_metrics_from_counts` turns the counters
into exactly what `compute_node_metrics` returns.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _cabi
from .topology import topology_for


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class LinkMetrics:
    def __init__(self, edge_index_routes: torch.Tensor, n_links: int, replicas: int = 1, n_hours: int = 25,
                 optimality: bool = True, device=None):
        dev = torch.device(device) if device is not None else edge_index_routes.device
        if dev.type != "cuda":
            raise RuntimeError("LinkMetrics lives on a CUDA device (no CPU fallback)")
        self.device, self.N, self.R, self.H = dev, int(n_links), int(replicas), int(n_hours)
        self.topo = topology_for(edge_index_routes, self.N)
        self.counts = torch.zeros(self.R, self.H, self.N, dtype=torch.int32, device=dev)
        self.optimality_sum = torch.zeros(self.R, self.H, self.N, dtype=torch.float32, device=dev) if optimality else None
        self.optimality_now = torch.zeros(self.R, self.N, dtype=torch.float32, device=dev) if optimality else None
        self.max_hour = -1        # highest hour recorded so far (the reference's num_hours = max_hour + 1)
        self.steps = 0

    def reset(self):
        self.counts.zero_()
        if self.optimality_sum is not None:
            self.optimality_sum.zero_()
            self.optimality_now.zero_()
        self.max_hour, self.steps = -1, 0

    def _grow(self, hour: int):
        H = max(hour + 1, 2 * self.H)
        for name in ("counts", "optimality_sum"):
            old = getattr(self, name)
            if old is not None:
                new = torch.zeros(self.R, H, self.N, dtype=old.dtype, device=self.device)
                new[:, : self.H] = old
                setattr(self, name, new)
        self.H = H

    @staticmethod
    def hour_of(time) -> int:
        """hours = (times // 3600).clamp(min=0) on a torch.long tensor built from the recorded times
        (src/transportation_simulator.py:596-602): the time is truncated to an integer first."""
        return max(int(time) // 3600, 0)

    def record(self, time, pop: torch.Tensor | None = None, withdrawn: torch.Tensor | None = None,
               delta_tt: torch.Tensor | None = None):
        """One step's side outputs: pop / withdrawn are bool or uint8 [R, N] (or [N] when R == 1), delta_tt fp32
        [R, E] (or [E]) in original edge order. Asynchronous on the current stream."""
        hour = self.hour_of(time)
        if hour >= self.H:
            self._grow(hour)

        def mask_ptr(m):
            if m is None:
                return None
            if m.dtype not in (torch.bool, torch.uint8) or m.numel() != self.R * self.N or m.device != self.device:
                raise ValueError("masks must be bool/uint8 with one entry per (replica, link) on the metrics device")
            if not m.is_contiguous():
                m = m.contiguous()
            keep.append(m)
            return m.data_ptr()

        keep = []
        pp, wp = mask_ptr(pop), mask_ptr(withdrawn)
        dp = None
        if delta_tt is not None and self.optimality_now is not None:
            if delta_tt.dtype != torch.float32 or delta_tt.numel() != self.R * self.topo.n_edges or delta_tt.device != self.device:
                raise ValueError("delta_tt must be fp32 with one entry per (replica, dual edge)")
            delta_tt = delta_tt.contiguous()
            keep.append(delta_tt)
            dp = delta_tt.data_ptr()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_metrics_accumulate(
                self.topo.ref(), self.R, pp, wp, dp, hour, self.H, self.counts.data_ptr(),
                self.optimality_sum.data_ptr() if dp is not None else None,
                self.optimality_now.data_ptr() if dp is not None else None, _stream(self.device))
        _cabi.check(rc, "tarl_metrics_accumulate")
        self.max_hour = max(self.max_hour, hour)
        self.steps += 1

    def counts_per_node(self, replica: int = 0) -> torch.Tensor:
        """int64 [N, num_hours] — `counts_per_node` of src/transportation_simulator.py:610-613."""
        return self.counts[replica, : self.max_hour + 1].t().to(torch.int64)


def counts_from_histories(update_history, withdraw_history, device=None):
    """The reference's own reduction (src/transportation_simulator.py:584-613) of the (time, bool[N]) histories:
    int64 [N, num_hours], or None when both are empty. Used when no on-device counters were kept."""
    combined = list(update_history) + list(withdraw_history)
    if not combined:
        return None
    dev = device if device is not None else combined[0][1].device
    hours = torch.tensor([LinkMetrics.hour_of(t) for t, _ in combined], dtype=torch.long, device=dev)
    masks = torch.stack([m.reshape(-1).to(dev) for _, m in combined], dim=0).to(torch.long)          # (T, N)
    num_hours = int(hours.max().item()) + 1
    out = torch.zeros(num_hours, masks.size(1), dtype=torch.long, device=dev)
    out.index_add_(0, hours, masks)
    return out.t().contiguous()


def node_metrics_from_counts(counts_per_node: torch.Tensor, max_flow: torch.Tensor, output_dir: str | None = None):
    """V/C statistics and the return value / CSV of compute_node_metrics (src/transportation_simulator.py:615-669):
    vc = counts / capacity (capacity 0 -> NaN), avg_vc = nanmean over hours, std_vc = population std over hours."""
    num_nodes, num_hours = counts_per_node.shape
    cap_safe = max_flow[:num_nodes].to(torch.float32).clone()
    cap_safe[cap_safe == 0] = float("nan")
    vc = counts_per_node.float() / cap_safe.unsqueeze(1)
    avg_np = torch.nanmean(vc, dim=1).cpu().numpy()
    std_np = torch.std(vc, dim=1, unbiased=False).cpu().numpy()
    counts_np = counts_per_node.cpu().numpy()
    if output_dir is not None:
        import pandas as pd
        df = pd.DataFrame(counts_np, columns=[f"count_{h}h" for h in range(num_hours)])
        df["node_id"] = range(num_nodes)
        df["avg_vc"] = avg_np
        df["std_vc"] = std_np
        df = df[["node_id", "avg_vc", "std_vc"] + [f"count_{h}h" for h in range(num_hours)]]
        os.makedirs(output_dir, exist_ok=True)
        df.to_csv(os.path.join(output_dir, "node_metrics.csv"), index=False)
        print(f"Wrote {os.path.join(output_dir, 'node_metrics.csv')}")
    return {n: {"avg_vc": float(avg_np[n]), "std_vc": float(std_np[n]), "hourly_counts": counts_np[n].tolist()}
            for n in range(num_nodes)}
