"""Auxiliary subsystems the reference lacks: device timers, NVTX
ranges, clock sampling, checkpoint/resume, structured logging."""
from .timers import DeviceTimer, PhaseTimers, nvtx_range, ClockSampler, l2_flush  # noqa: F401
from .checkpoint import save_checkpoint, load_checkpoint  # noqa: F401
