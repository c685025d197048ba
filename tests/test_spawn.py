"""Script launcher (dist_tuto.pth_b200.spawn): env rendezvous, failure propagation, timeout -- the supervised
counterpart of the reference's `__main__` fork/join block (train_dist.py:138-147)."""
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "spawn_scripts.py")
pytestmark = pytest.mark.timeout(240)


def launch(*args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    t0 = time.monotonic()
    p = subprocess.run([sys.executable, "-m", "dist_tuto.pth_b200.spawn", *args], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=200)
    return p, time.monotonic() - t0


def test_ranks_rendezvous_through_the_env():
    p, _ = launch("--size", "3", SCRIPT, "allreduce")
    assert p.returncode == 0, p.stdout + p.stderr
    lines = sorted(l for l in p.stdout.splitlines() if l.startswith("rank"))
    assert lines == [f"rank {r} of 3 sum 6.0" for r in range(3)]


def test_first_failure_stops_the_job_with_that_exit_code():
    p, dt = launch("--size", "3", SCRIPT, "fail")
    assert p.returncode == 7 and "rank 1 exited with code 7" in p.stderr
    assert dt < 60                                       # the sleeping survivors were terminated, not joined


def test_timeout_kills_every_rank():
    p, dt = launch("--size", "2", "--timeout", "2", SCRIPT, "hang")
    assert p.returncode == 124 and "exceeded" in p.stderr and dt < 60


def test_run_script_api_returns_zero(tmp_path):
    from dist_tuto.pth_b200.spawn import run_script
    ok = tmp_path / "ok.py"
    ok.write_text("import os, sys\nassert int(os.environ['WORLD_SIZE']) == 2 and os.environ['MASTER_ADDR'] == '127.0.0.1'\n"
                  "assert os.environ['RANK'] == os.environ['LOCAL_RANK']\n")
    assert run_script(str(ok), size=2, timeout_s=60) == 0


def test_two_node_job_on_localhost():
    """Two launchers = two 'machines' (tuto.md:404-428): 2 x 2 ranks rendezvous at one master, global ranks 0..3."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = str(s.getsockname()[1])
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    nodes = [subprocess.Popen([sys.executable, "-m", "dist_tuto.pth_b200.spawn", "--size", "2", "--nnodes", "2", "--node-rank", str(n),
                               "--master-port", port, "--timeout", "120", SCRIPT, "allreduce"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT) for n in (0, 1)]
    outs = [p.communicate(timeout=200) for p in nodes]
    assert [p.returncode for p in nodes] == [0, 0], outs
    lines = sorted(l for o, _ in outs for l in o.splitlines() if l.startswith("rank"))
    assert lines == [f"rank {r} of 4 sum 10.0" for r in range(4)]


def test_multi_node_needs_an_explicit_port():
    from dist_tuto.pth_b200.spawn import run_script
    with pytest.raises(ValueError, match="master_port"):
        run_script(SCRIPT, ["allreduce"], size=1, nnodes=2, node_rank=0)


def test_restart_after_a_failure_resumes_from_the_last_checkpoint(tmp_path):
    """--max-restarts: rank 0 dies at the end of epoch 1 of 3 on the first attempt (before checkpointing it); the second attempt
    resumes from the checkpoint written after epoch 0 (TrainConfig.checkpoint_every) and runs exactly the two remaining epochs."""
    import torch
    ckpt = str(tmp_path / "job.pt")
    p, _ = launch("--size", "2", "--max-restarts", "1", SCRIPT, "train_restart", ckpt)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "rank 0 exited with code 9" in p.stderr and "restarting the job (attempt 2 of 2)" in p.stderr
    blob = torch.load(ckpt, map_location="cpu")
    nb = 512 // 2 // 64                                     # 2 ranks x batch 64: 4 steps per epoch
    assert blob["epoch"] == 3 and blob["in_progress"] is False and len(blob["history"]) == 3 and blob["steps"] == 3 * nb
    notes = open(ckpt + ".epochs").read().strip().splitlines()
    assert notes == ["attempt 1: epochs [1, 2] history 3"], notes          # attempt 0 never finished; attempt 1 did epochs 1 and 2
    # without restarts the same failure ends the job with the rank's exit code
    p2, _ = launch("--size", "2", SCRIPT, "train_restart", str(tmp_path / "job2.pt"))
    assert p2.returncode == 9
