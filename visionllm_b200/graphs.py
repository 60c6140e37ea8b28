"""CUDA-graph capture of a launch-bound forward (the Grounding-DINO stage issues ~1700 small launches per step and is
host-bound when launched eagerly from Python).  One graph per input signature; replays copy the inputs into the
graph's static buffers and return the graph's static outputs (valid until the next replay of the same signature).

The captured function must be free of host synchronisation (no .item() / .tolist() / pageable H2D copies): the GDINO
path keeps host copies of its integer shape tensors (`msda.attach_host_shapes`) for exactly this reason.  Our kernels
are captured like any other launch: the ctypes wrappers enqueue on torch's current stream, which is the capture stream
inside `torch.cuda.graph`, and TMA descriptors are kernel parameters, so they are baked into the graph nodes."""
import torch

from . import _lib, ops


class GraphedForward:
    def __init__(self, fn, warmup=2):
        self.fn, self.warmup = fn, warmup
        self._cache = {}
        self.launches_per_replay = 0

    @staticmethod
    def _sig(args):
        return tuple(None if a is None else (tuple(a.shape), a.dtype, a.device.index) for a in args)

    def _capture(self, args):
        static_in = [None if a is None else a.clone() for a in args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):                     # allocates workspaces / packs weights / builds index caches
                self.fn(*static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            out = self.fn(*static_in)
        self.launches_per_replay = _lib.launch_count() - n0
        keep = (dict(ops._ATTN_WS), dict(ops._GN_WS))       # workspaces the graph's nodes point into stay alive
        return static_in, graph, out, keep

    def __call__(self, *args):
        sig = self._sig(args)
        entry = self._cache.get(sig)
        if entry is None:
            entry = self._cache[sig] = self._capture(args)
        static_in, graph, out, _ = entry
        for s, a in zip(static_in, args):
            if s is not None and s.data_ptr() != a.data_ptr():
                s.copy_(a, non_blocking=True)
        graph.replay()
        _lib.add_launches(self.launches_per_replay)
        return out
