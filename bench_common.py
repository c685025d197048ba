"""Shared, framework-free helpers for bench.py (both arms): clock sampling, max-over-ranks, the JSON
result line, and synthetic MNIST idx files for the reference's unmodified ``partition_dataset()``.
No model / kernel / engine code lives here."""
from __future__ import annotations

import json
import os
import statistics
import struct
import subprocess
import tempfile
import threading

METRIC = "MNIST-ConvNet samples/sec (whole box, device-timed, max over ranks)"
_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi SM clock + throttle-reason sampler running during the timed region."""

    def __init__(self, gpu_index=0, period_s=0.1):
        if isinstance(gpu_index, int):      # logical CUDA index -> physical GPU (CUDA_VISIBLE_DEVICES re-numbers devices)
            try:
                import torch
                u = str(torch.cuda.get_device_properties(gpu_index).uuid)
                gpu_index = u if u.startswith("GPU-") else "GPU-" + u
            except Exception:
                pass
        self.gpu, self.period, self.rows = gpu_index, period_s, []
        self._stop, self._th = threading.Event(), None

    def _once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={_Q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(self.period)

    def __enter__(self):
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=6)
        if not self.rows:
            self._once()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def max_over_ranks(value, device):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aligned_start(device, lead_s=0.003):
    """Barrier + synchronize, then every rank leaves at the same instant; returns that instant (``time.perf_counter()``).

    A barrier alone lets the ranks go as each host thread happens to notice it (NCCL completion polling, Python): a skew
    of 100+ us, which a 20-step window of ~30 us steps then reports as "step time" because the first gradient exchange
    waits for the last rank to start.  All ranks of this single-node job share CLOCK_MONOTONIC (what perf_counter reads on
    Linux), so they agree on a start time a few milliseconds ahead (MAX over ranks) and spin until it.  Used by both arms."""
    import time
    import torch
    import torch.distributed as dist
    multi = dist.is_initialized() and dist.get_world_size() > 1
    on_gpu = torch.cuda.is_available() and torch.device(device).type == "cuda"
    if multi:
        dist.barrier()
    if on_gpu:
        torch.cuda.synchronize()
    if not multi:
        return time.perf_counter()
    t = torch.tensor([time.perf_counter() + lead_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    start = float(t.item())
    if on_gpu:
        torch.cuda.synchronize()
    while time.perf_counter() < start:
        pass
    return max(start, time.perf_counter())


def result_line(impl, value, ms, n_gpus, steps, warmup, clocks, e2e_value, h2d, d2h, gpu_launches, dtype,
                extra_config=None):
    cfg = {"model": "MNIST ConvNet (train_dist.py Net, 21,840 params)", "global_batch": 128, "seq_len": None,
           "image": "1x28x28", "parallelism": f"dp{n_gpus}", "optimizer": "SGD lr=0.01 momentum=0.5",
           "per_gpu_batch": 128 // n_gpus}
    cfg.update(extra_config or {})
    out = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / max(steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": dtype, "data": "synthetic 28x28 images, random-init weights", "config": cfg, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": gpu_launches, "impl": impl}
    return json.dumps(out)


def synthetic_idx_dir(rank=0, n=60000, seed=1234):
    """A scratch cwd whose ./data/MNIST/raw holds synthetic idx files (MNIST layout, random content)."""
    import numpy as np
    base = os.path.join(tempfile.gettempdir(), f"b2_ref_data_{os.getuid()}")
    raw = os.path.join(base, "data", "MNIST", "raw")
    marker = os.path.join(raw, ".complete")
    if not os.path.isfile(marker):
        if rank == 0 or not os.path.isdir(raw):
            os.makedirs(raw, exist_ok=True)
            rng = np.random.default_rng(seed)

            def dump(img, lab, count):
                with open(os.path.join(raw, img + f".tmp{os.getpid()}"), "wb") as f:
                    f.write(struct.pack(">IIII", 2051, count, 28, 28))
                    f.write(rng.integers(0, 256, size=(count, 28, 28), dtype=np.uint8).tobytes())
                os.replace(os.path.join(raw, img + f".tmp{os.getpid()}"), os.path.join(raw, img))
                with open(os.path.join(raw, lab + f".tmp{os.getpid()}"), "wb") as f:
                    f.write(struct.pack(">II", 2049, count))
                    f.write(rng.integers(0, 10, size=(count,), dtype=np.uint8).tobytes())
                os.replace(os.path.join(raw, lab + f".tmp{os.getpid()}"), os.path.join(raw, lab))

            dump("train-images-idx3-ubyte", "train-labels-idx1-ubyte", n)
            dump("t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte", 1000)
            open(marker, "w").write("ok")
    import time
    for _ in range(600):
        if os.path.isfile(marker):
            break
        time.sleep(0.1)
    return base
