"""CUDA-graph replay of a whole SR forward (SwinIR / HAT / DAT drop-in models).

A forward is several hundred kernel launches (36 fused blocks x 2-8 kernels + cuDNN convolutions + the tail) of 10-100 us
each, so at batch 8-16 the host's launch rate, not the GPU, bounds the step.  ``GraphedModel`` captures one forward per
input shape into a ``torch.cuda.CUDAGraph`` -- libsrk's entry points only enqueue on the current stream and never
allocate or synchronise, so they capture like any other kernel -- and replays it with one ``cudaGraphLaunch``.

    g = GraphedModel(model)          # model: srk.SwinIR / srk.HAT, eval mode, on the GPU
    y = g(x)                         # first call per shape: warm-up + capture; later calls: copy-in, replay.
                                     # x may be a (pinned) host tensor: it is copied straight into the graph's input buffer

The returned tensor is the graph's static output buffer: it is overwritten by the next call with the same shape
(``clone()`` it to keep it).  Weights are baked in by address: after changing parameters call ``reset()``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn


class GraphedModel(nn.Module):
    def __init__(self, model: nn.Module, warmup: int = 2):
        super().__init__()
        self.model = model
        self.warmup = warmup
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor, torch.Tensor]] = {}
        self.h2d_done = None            # event recorded after the last host -> device input copy (None: the last input was on the device)

    def reset(self) -> None:
        self._graphs.clear()

    @torch.no_grad()
    def _capture(self, x: torch.Tensor):
        static_in = x.clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture: packs weights, lets cuDNN pick algorithms
            for _ in range(self.warmup):
                self.model(static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self.model(static_in)
        return graph, static_in, static_out

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedModel: the model must live on a CUDA device (no CPU fallback)")
        key = (tuple(x.shape), x.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(x.to(dev))
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        static_in.copy_(x, non_blocking=True)       # any strides / memory format of x: copy_ converts into the captured buffer's layout
        if not x.is_cuda:
            # a pinned host source is read asynchronously: it may be refilled only after this event (PipelinedRunner.submit returns it)
            self.h2d_done = torch.cuda.Event()
            self.h2d_done.record(torch.cuda.current_stream(dev))
        graph.replay()
        return static_out


class PipelinedRunner:
    """Host-to-host serving loop around a ``GraphedModel``: ``submit(host_in, host_out)`` copies a pinned host batch into the
    graph's input buffer, replays the forward and copies the result into the pinned ``host_out`` -- the device-to-host copy
    runs on a side stream, so it overlaps the next batch's forward (the 12.6 MB SR batch of SwinIR x4 B=16 takes ~0.25 ms of
    PCIe time, 5 % of a step).  Results are complete after ``drain()`` (or once ``submit`` has been called ``depth`` more
    times).  ``depth`` device-side output buffers decouple the graph's static output from the copies in flight.
    """

    def __init__(self, graphed: GraphedModel, depth: int = 2):
        self.graphed = graphed
        self.depth = depth
        self._copy_stream = None
        self._bufs = []          # device staging buffers
        self._ready = []         # forward finished -> staging buffer holds the result
        self._done = []          # D2H finished -> staging buffer free
        self._i = 0

    @torch.no_grad()
    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor):
        """Returns the event after which `host_in` has been read (None for a device input): a serving loop that reuses one pinned
        input buffer must `event.synchronize()` before refilling it -- the H2D copy is asynchronous, and only the D2H side is
        covered by `drain()`.  (Rotating `depth + 1` input buffers, as bench.py does, needs no wait.)"""
        y = self.graphed(host_in)                                   # H2D into the static input + replay, current stream
        h2d_done = self.graphed.h2d_done if not host_in.is_cuda else None
        cur = torch.cuda.current_stream(y.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=y.device)
            self._bufs = [torch.empty_like(y) for _ in range(self.depth)]
            self._ready = [torch.cuda.Event() for _ in range(self.depth)]
            self._done = [None] * self.depth
        b = self._i % self.depth
        self._i += 1
        if self._done[b] is not None:
            cur.wait_event(self._done[b])                           # the copy that last read this staging buffer has finished
        self._bufs[b].copy_(y, non_blocking=True)                   # device-to-device: frees the graph's static output
        self._ready[b].record(cur)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ready[b])
            host_out.copy_(self._bufs[b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            self._done[b] = ev
        return h2d_done

    def drain(self) -> None:
        """Make the current stream wait for every copy in flight (then synchronise the stream / an event as usual)."""
        if self._copy_stream is not None:
            torch.cuda.current_stream(self._bufs[0].device).wait_stream(self._copy_stream)

