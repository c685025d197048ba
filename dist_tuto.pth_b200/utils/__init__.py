"""Auxiliary subsystems the reference lacks: device timers, NVTX
ranges, clock sampling, checkpoint/resume, structured logging."""
from .timers import DeviceTimer, PhaseTimers, nvtx_range, ClockSampler, l2_flush  # noqa: F401
from .checkpoint import save_checkpoint, load_checkpoint  # noqa: F401


def say(*args, sep=" ", end="\n"):
    """``print(*args)`` as a single write + flush: the ranks of a job share one terminal or pipe, and the tutorial's
    multi-argument prints (tuto.md:91 ``print('Rank ', rank, ' has data ', tensor[0])``) would interleave mid-line."""
    import sys
    sys.stdout.write(sep.join(str(a) for a in args) + end)
    sys.stdout.flush()
