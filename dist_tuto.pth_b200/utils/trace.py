"""Timeline tracing: host spans + CUDA-event spans of every rank in ONE Chrome / Perfetto trace file.

The reference has no tracing at all.  ``Tracer`` records

* host spans   -- ``with tracer.span("load batch"):`` (``perf_counter_ns``), and
* device spans -- ``with tracer.device_span("step kernels", stream):`` two CUDA events on the launching stream; they are
  turned into timestamps at ``save()`` by ``elapsed_time`` against an anchor event whose host time is known, so recording
  costs two ``cudaEventRecord`` and never synchronises inside the traced region,

and ``save(path)`` gathers the events of all ranks on rank 0 and writes the Trace Event Format (``chrome://tracing``,
https://ui.perfetto.dev): one process row per rank, one thread row for the host and one per traced stream.  All ranks of a
single-node job share CLOCK_MONOTONIC, so their rows line up without any clock exchange.

    tracer = Tracer()                      # or TrainConfig(trace="run.trace.json"): train() does the rest
    with tracer.span("epoch 0"):
        ...
    tracer.save("run.trace.json")
"""
from __future__ import annotations

import contextlib
import json
import os
import time
from typing import Any, Dict, List, Optional

import torch

__all__ = ["Tracer", "NullTracer"]


class NullTracer:
    """Same surface as :class:`Tracer`, records nothing (the default inside ``train()``)."""

    enabled = False

    @contextlib.contextmanager
    def span(self, name: str, cat: str = "host", **args):
        yield

    @contextlib.contextmanager
    def device_span(self, name: str, stream=None, cat: str = "gpu", **args):
        yield

    def instant(self, name: str, **args):
        pass

    def counter(self, name: str, value: float):
        pass

    def save(self, path: str, group=None) -> Optional[str]:
        return None


class Tracer(NullTracer):
    """Collects spans of this rank; see the module docstring."""

    enabled = True

    def __init__(self, rank: Optional[int] = None, max_events: int = 200000):
        import torch.distributed as dist
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_available() and dist.is_initialized() else 0)
        self.max_events = max_events
        self._host: List[Dict[str, Any]] = []
        self._dev: List[Any] = []                 # (name, cat, args, stream id, e0, e1)
        self._anchor = None                       # (cuda event, host ns when it was recorded and known complete)
        self.dropped = 0

    # ------------------------------------------------------------------ recording
    def _room(self) -> bool:
        if len(self._host) + len(self._dev) >= self.max_events:
            self.dropped += 1
            return False
        return True

    @contextlib.contextmanager
    def span(self, name: str, cat: str = "host", **args):
        t0 = time.perf_counter_ns()
        try:
            yield
        finally:
            if self._room():
                self._host.append({"name": name, "cat": cat, "ph": "X", "ts": t0 / 1e3, "dur": (time.perf_counter_ns() - t0) / 1e3,
                                   "pid": self.rank, "tid": "host", "args": args})

    def instant(self, name: str, **args):
        if self._room():
            self._host.append({"name": name, "ph": "i", "s": "p", "ts": time.perf_counter_ns() / 1e3, "pid": self.rank, "tid": "host",
                               "args": args})

    def counter(self, name: str, value: float):
        if self._room():
            self._host.append({"name": name, "ph": "C", "ts": time.perf_counter_ns() / 1e3, "pid": self.rank, "args": {name: value}})

    def _ensure_anchor(self):
        if self._anchor is None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            ev.synchronize()                      # once, outside any traced region: host time of a completed event
            self._anchor = (ev, time.perf_counter_ns())

    @contextlib.contextmanager
    def device_span(self, name: str, stream=None, cat: str = "gpu", **args):
        if not torch.cuda.is_available():
            with self.span(name, cat, **args):    # CPU build: degrade to a host span
                yield
            return
        self._ensure_anchor()
        st = stream if stream is not None else torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        try:
            yield
        finally:
            e1.record(st)
            if self._room():
                self._dev.append((name, cat, args, int(st.cuda_stream), e0, e1))

    # ------------------------------------------------------------------ export
    def events(self) -> List[Dict[str, Any]]:
        """This rank's events in Trace Event Format (synchronises the traced streams' last events)."""
        out = list(self._host)
        if self._dev:
            anchor_ev, anchor_ns = self._anchor
            for name, cat, args, sid, e0, e1 in self._dev:
                e1.synchronize()
                t0 = anchor_ns / 1e3 + anchor_ev.elapsed_time(e0) * 1e3          # us
                out.append({"name": name, "cat": cat, "ph": "X", "ts": t0, "dur": max(e0.elapsed_time(e1) * 1e3, 0.001),
                            "pid": self.rank, "tid": f"stream {sid:#x}", "args": args})
        return out

    def save(self, path: str, group=None) -> Optional[str]:
        """Collective when a process group is initialised: rank 0 writes every rank's events to ``path``; returns the path
        on the writing rank, ``None`` elsewhere."""
        import torch.distributed as dist
        mine = self.events()
        meta = [{"name": "process_name", "ph": "M", "pid": self.rank, "args": {"name": f"rank {self.rank}"}}]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            allev: List[Any] = [None] * dist.get_world_size(group)
            dist.all_gather_object(allev, meta + mine, group=group)
            writer = dist.get_global_rank(group, 0) if group is not None else 0
            if dist.get_rank() != writer:
                return None
            events = [e for part in allev for e in part]
        else:
            events = meta + mine
        tmp = f"{path}.tmp.{os.getpid()}"
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(tmp, "w") as f:
            json.dump({"traceEvents": events, "displayTimeUnit": "ms",
                       "otherData": {"library": "dist_tuto.pth_b200", "dropped_events": self.dropped}}, f)
        os.replace(tmp, path)
        return path
