"""The tutorial's demo programs (reference: ptp.py, gloo.py/allreduce.py, tuto.md snippets, train_dist.py __main__) run
as real processes on CPU/gloo and print what the tutorial says they print."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(300)


def run(script, *args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True,
                       timeout=280, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_p2p_demo_prints_one_on_both_ranks():
    out = run("p2p_demo.py")                                 # tuto.md:91,116: "Rank 1 has data 1.0"
    assert re.search(r"\[blocking\]\s+Rank\s+1\s+has data\s+1\.0", out)
    assert len(re.findall(r"\[non-blocking\] Rank\s+\d\s+has data\s+1\.0", out)) == 2


def test_gather_demo_rank0_prints_world_sum():
    out = run("gather_demo.py")                              # ptp.py:28 prints the sum of the gathered ones
    assert re.search(r"^22$|^2$|^\d+$", out, re.M)


def test_groups_demo_prints_two():
    out = run("groups_demo.py")                              # tuto.md:185
    assert len(re.findall(r"Rank\s+\d\s+has data\s+2\.0", out)) == 2


def test_allreduce_demo_ring_equals_collective():
    out = run("allreduce_demo.py", "--size", "3")            # gloo.py: ring allreduce vs dist.all_reduce
    rings = re.findall(r"ring:\s+(\[.*?\])", out)
    colls = re.findall(r"collective:\s+(\[.*?\])", out)
    assert len(rings) == 3 and len(colls) == 3 and len(set(rings + colls)) == 1


def test_train_mnist_example_two_ranks_few_steps():
    out = run("train_mnist.py", "--size", "2", "--backend", "gloo", "--epochs", "1", "--max-steps", "6")
    lines = re.findall(r"Rank\s+(\d)\s*, epoch\s+0\s*:\s+([0-9.]+)", out)     # train_dist.py:125-127 print format
    assert sorted(r for r, _ in lines) == ["0", "1"]
    assert all(0.5 < float(v) < 5.0 for _, v in lines)


def torchrun(nproc, script, *args):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "examples", script), *args],
                       capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_external_launcher_env_recipe_train():
    """rank/size from an external launcher (the tutorial's MPI recipe, tuto.md:383-398; here torchrun provides the env)."""
    out = torchrun(2, "train_mnist.py", "--external", "--backend", "gloo", "--epochs", "1", "--max-steps", "4")
    assert sorted(re.findall(r"Rank\s+(\d)\s*, epoch\s+0", out)) == ["0", "1"]


def test_mpi_backend_recipe_allreduce_demo():
    out = torchrun(3, "allreduce_demo.py", "--backend", "mpi")          # init_processes(0, 0, run, backend='mpi')
    rings = re.findall(r"ring:\s+(\[.*?\])", out)
    assert len(rings) == 3 and len(set(rings)) == 1
