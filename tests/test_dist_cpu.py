"""Distributed-CPU tier: N local processes on gloo @127.0.0.1."""

import pytest

import dist_tuto.pth_b200 as b2
import dist_workers as W

pytestmark = pytest.mark.timeout(240)


def go(fn, size, backend="gloo", **kw):
    b2.launch(fn, size=size, backend=backend, join_timeout_s=200, **kw)


def test_p2p_blocking_and_nonblocking():
    go(W.w_p2p, 2)


def test_collectives_ops_groups_world3():
    go(W.w_collectives, 3)


def test_ring_allreduce_rank_dependent_world3():
    go(W.w_ring, 3)


def test_ring_allreduce_world2():
    go(W.w_ring, 2)


def test_average_gradients_and_ddp_world2():
    go(W.w_average_gradients, 2)


def test_flat_sgd_with_ddp_world2_equals_global_batch_sgd():
    go(W.w_flat_sgd_ddp, 2)


def test_symmetric_memory_fd_exchange_world3():
    go(W.w_symm_fd_exchange, 3)


def test_train_loop_world2_replicas_identical():
    go(W.w_train, 2)


def test_tcp_backend_maps_to_gloo():
    with pytest.warns(None) if False else _nullctx():
        go(W.w_env, 2, backend="tcp")


def test_file_init_method(tmp_path):
    go(W.w_env, 2, init_method=f"file://{tmp_path}/rdzv")


def test_tcp_init_method():
    port = b2.find_free_port()
    go(W.w_env, 2, init_method=f"tcp://127.0.0.1:{port}")


def test_failure_propagates_and_survivors_are_killed():
    with pytest.raises(b2.LaunchError) as ei:
        b2.launch(W.w_fail, size=2, backend="gloo", join_timeout_s=120)
    assert ei.value.rank == 1 and "boom on rank 1" in ei.value.child_traceback


def test_join_timeout():
    with pytest.raises(b2.LaunchError) as ei:
        b2.launch(W.w_hang, size=2, backend="gloo", join_timeout_s=8)
    assert ei.value.exitcode == "timeout"


def test_external_launcher_env_mpi_recipe(monkeypatch):
    # tuto.md:393-398: init_processes(0, 0, run, backend='mpi'); rank/size from the launcher env
    port = b2.find_free_port()
    monkeypatch.setenv("OMPI_COMM_WORLD_RANK", "0")
    monkeypatch.setenv("OMPI_COMM_WORLD_SIZE", "1")
    monkeypatch.setenv("MASTER_PORT", str(port))
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    seen = {}
    b2.init_processes(0, 0, lambda r, s: seen.update(r=r, s=s, ws=b2.get_world_size()), backend="mpi")
    assert seen == {"r": 0, "s": 1, "ws": 1}
    assert not b2.is_initialized()            # torn down (D8)


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_file_init_method_with_group_name(tmp_path):
    go(W.w_env, 2, init_method=f"file://{tmp_path}/shared", group_name="job-a")     # tuto.md:437-448


def test_multicast_and_unknown_init_methods_are_rejected_with_a_clear_error():
    import importlib
    L = importlib.import_module("dist_tuto.pth_b200.launch")     # (the package attribute `launch` is the function)
    with pytest.raises(ValueError, match="multicast"):
        L._init_method("tcp://[ff15:1e18:5d4c:4cf0:d02d:b659:53ba:b0a7]:23456", "127.0.0.1", 29500)   # tuto.md:452
    with pytest.raises(ValueError, match="unsupported"):
        L._init_method("zeromq://x", "127.0.0.1", 29500)
    assert L._init_method("tcp://10.1.1.20:23456", "127.0.0.1", 29500) == "tcp://10.1.1.20:23456"       # tuto.md:432


def test_mpi_backend_without_a_launcher_environment_fails_loudly(monkeypatch):
    import importlib
    L = importlib.import_module("dist_tuto.pth_b200.launch")     # (the package attribute `launch` is the function)
    for k in ("RANK", "WORLD_SIZE", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE", "SLURM_PROCID",
              "SLURM_NTASKS"):
        monkeypatch.delenv(k, raising=False)
    with pytest.raises(RuntimeError, match="external launcher"):
        L.init_processes(0, 0, lambda r, s: None, backend="mpi")


def test_symmetric_backend_refuses_to_span_machines():
    go(W.w_one_node_guard, 2)


def test_two_level_world_over_simulated_machines():
    """A multi-machine b200 job gets parallel/hier.HierWorld: peer memory inside a machine, one NCCL rail per local rank
    across machines.  Here: gloo, 4 ranks on 2 simulated machines."""
    go(W.w_hier_world, 4)


def test_train_writes_one_trace_for_all_ranks():
    go(W.w_train_trace, 2)
