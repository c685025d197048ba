"""Script launcher: ``python -m dist_tuto.pth_b200.spawn --size 8 script.py [args...]``.

The reference starts its ranks from inside each script (``train_dist.py:138-147``: a literal world size, ``Process`` per
rank, ``join()`` forever) or hands the job to ``mpirun`` (tuto.md:396).  This is the production counterpart of that
``__main__`` block for scripts that call :func:`dist_tuto.pth_b200.init_from_env` (the "mpi" recipe): it starts one
process per rank (per GPU on the ``b200`` backend) with ``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` / ``MASTER_ADDR`` /
``MASTER_PORT`` set, forwards their output, and -- unlike the reference -- supervises them:

* the first rank that exits non-zero ends the job: the survivors are terminated (they would otherwise hang in a
  collective, SURVEY §5 "failure detection") and the launcher exits with that rank's code;
* ``--max-restarts N``: after a failure the WHOLE job is started again (fresh rendezvous port, ``B200DIST_RESTART_COUNT`` =
  1, 2, ... in the environment) up to N times -- a script that checkpoints (``TrainConfig(checkpoint=..., checkpoint_every=1)``)
  and resumes when the file exists continues where the last checkpoint left off (``examples/train_mnist.py`` does);
* ``--timeout`` bounds the whole job; ``SIGINT`` / ``SIGTERM`` are forwarded to every rank;
* only the processes started here are ever signalled (exact PIDs).
"""
from __future__ import annotations

import argparse
import os
import signal
import subprocess
import sys
import time
from typing import List, Optional, Sequence

from .launch import DEFAULT_ADDR, find_free_port

__all__ = ["run_script", "main"]


def _stop(procs: Sequence[subprocess.Popen], grace_s: float = 5.0) -> None:
    for p in procs:
        if p.poll() is None:
            p.terminate()
    t_end = time.monotonic() + grace_s
    for p in procs:
        while p.poll() is None and time.monotonic() < t_end:
            time.sleep(0.05)
        if p.poll() is None:
            p.kill()
    for p in procs:
        p.wait()


def run_script(script: str, script_args: Sequence[str] = (), size: int = 2, master_addr: str = DEFAULT_ADDR,
               master_port: Optional[int] = None, timeout_s: Optional[float] = None, env: Optional[dict] = None,
               poll_s: float = 0.1, nnodes: int = 1, node_rank: int = 0, max_restarts: int = 0) -> int:
    """Run ``script`` as ``size`` local ranks; returns the job's exit code (0 = every rank exited 0).

    Multi-machine jobs (the tutorial's "replace MASTER_ADDR by the IP of the master", tuto.md:404-428): start the launcher
    once per machine with the same ``master_addr``/``master_port`` and ``nnodes``, and ``node_rank`` = 0..nnodes-1; global
    rank = ``node_rank * size + local rank``.  (Across machines the ``b200`` backend builds its two-level world,
    parallel/hier.py: peer memory inside each machine, NCCL between them.)

    ``max_restarts``: how many times a failed job (a rank exited non-zero) is started over; the timeout covers all attempts.

    Exit codes: the failing rank's own code (of the last attempt); 124 on timeout (like ``timeout(1)``); 130 on interrupt."""
    if nnodes > 1 and master_port is None:
        raise ValueError("multi-node jobs need an explicit master_port (the same on every node)")
    if not 0 <= node_rank < nnodes:
        raise ValueError("node_rank must be in [0, nnodes)")
    deadline = None if timeout_s is None else time.monotonic() + timeout_s
    attempt = 0
    while True:
        code, failed = _run_once(script, script_args, size, master_addr, master_port, deadline, timeout_s, env, poll_s, nnodes,
                                 node_rank, attempt)
        if not failed or attempt >= max_restarts:
            return code
        attempt += 1
        sys.stderr.write(f"[dist_tuto.spawn] restarting the job (attempt {attempt + 1} of {max_restarts + 1})\n")


def _run_once(script, script_args, size, master_addr, master_port, deadline, timeout_s, env, poll_s, nnodes, node_rank, attempt):
    """One attempt; returns (exit code, restartable failure?)."""
    port = master_port or find_free_port(master_addr)
    base = dict(os.environ if env is None else env)
    base.update(WORLD_SIZE=str(size * nnodes), MASTER_ADDR=master_addr, MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(size),
                GROUP_RANK=str(node_rank), B200DIST_RESTART_COUNT=str(attempt))
    base.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, size))))
    procs: List[subprocess.Popen] = []
    interrupted = []
    old = {}

    def on_signal(signum, _frame):
        interrupted.append(signum)

    for sg in (signal.SIGINT, signal.SIGTERM):
        try:
            old[sg] = signal.signal(sg, on_signal)
        except ValueError:           # not the main thread
            pass
    try:
        for r in range(size):
            e = dict(base, RANK=str(node_rank * size + r), LOCAL_RANK=str(r))
            procs.append(subprocess.Popen([sys.executable, script, *script_args], env=e))
        while True:
            codes = [p.poll() for p in procs]
            bad = [(r, c) for r, c in enumerate(codes) if c not in (None, 0)]
            if bad:
                r, c = bad[0]
                sys.stderr.write(f"[dist_tuto.spawn] rank {r} exited with code {c}; stopping the other ranks\n")
                _stop(procs)
                return (c if c > 0 else 128 - c), True          # negative = killed by signal -c
            if all(c == 0 for c in codes):
                return 0, False
            if interrupted:
                sys.stderr.write("[dist_tuto.spawn] interrupted; stopping all ranks\n")
                _stop(procs)
                return 130, False
            if deadline is not None and time.monotonic() > deadline:
                sys.stderr.write(f"[dist_tuto.spawn] job exceeded {timeout_s:.0f} s; stopping all ranks\n")
                _stop(procs)
                return 124, False
            time.sleep(poll_s)
    finally:
        _stop([p for p in procs if p.poll() is None])
        for sg, h in old.items():
            signal.signal(sg, h)


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="python -m dist_tuto.pth_b200.spawn",
                                 description="start N ranks of a script that calls dist.init_from_env(run, backend)")
    ap.add_argument("--size", "-n", type=int, default=2, help="ranks on THIS machine (one process per rank / per GPU)")
    ap.add_argument("--master-addr", default=DEFAULT_ADDR)
    ap.add_argument("--master-port", type=int, default=None, help="default: a free port")
    ap.add_argument("--timeout", type=float, default=None, help="seconds for the whole job")
    ap.add_argument("--nnodes", type=int, default=1, help="machines in the job (run the launcher once per machine)")
    ap.add_argument("--node-rank", type=int, default=0, help="index of this machine, 0 = the one MASTER_ADDR points at")
    ap.add_argument("--max-restarts", type=int, default=0, help="start the whole job over after a failure, up to N times")
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    return run_script(a.script, a.script_args, size=a.size, master_addr=a.master_addr, master_port=a.master_port,
                      timeout_s=a.timeout, nnodes=a.nnodes, node_rank=a.node_rank, max_restarts=a.max_restarts)


if __name__ == "__main__":
    sys.exit(main())
