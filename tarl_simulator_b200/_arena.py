"""Chunked buffers for small per-step tensors that a history list retains."""
import torch

from . import _cabi


class StepArena:
    """Per-step outputs that the history retains (pop mask bool[N], flag words int32[4]) are handed out as rows of
    chunked buffers: retaining one small tensor per step straight from the caching allocator pins a fresh 20 MB
    segment every other step and turns every later torch.empty into a cudaMalloc (measured: 2 ms per step at 1M
    links); one allocation per `chunk` steps does not."""

    def __init__(self, chunk: int = 64):
        self.chunk, self.i, self.key = chunk, 0, None
        self.masks = self.flags = None

    def take(self, n_links: int, device):
        key = (n_links, device)
        if self.key != key or self.i >= self.chunk:
            self.masks = torch.empty(self.chunk, max(n_links, 1), dtype=torch.bool, device=device)
            self.flags = torch.zeros(self.chunk, _cabi.FLAG_COUNT, dtype=torch.int32, device=device)
            self.key, self.i = key, 0
        m, f = self.masks[self.i, :n_links], self.flags[self.i]
        self.i += 1
        return m, f
