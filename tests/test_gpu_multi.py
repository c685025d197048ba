"""GPU-multi tier: fused peer-memory all-reduce vs NCCL, DDP overlap, fused trainer, NCCL p2p ring."""
import pytest
import torch

import dist_tuto.pth_b200 as b2
import gpu_workers as W

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu, pytest.mark.timeout(900)]


def _n():
    n = torch.cuda.device_count()
    return 8 if n >= 8 else (4 if n >= 4 else 2)


def go(fn, size=None):
    b2.launch(fn, size=size or _n(), backend="b200", join_timeout_s=600)


def test_symmetric_allreduce_all_variants_vs_nccl():
    go(W.w_symm_allreduce)


def test_average_gradients_and_overlapped_ddp():
    go(W.w_average_gradients_gpu)


def test_fused_trainer_equals_global_batch_sgd():
    go(W.w_fused_trainer)


def test_nccl_p2p_and_ring_allreduce():
    go(W.w_p2p_ring_gpu)


def test_train_loop_fused_engine():
    go(W.w_train_fused_e2e)


def test_push_exchange_equals_barrier_exchange():
    go(W.w_push_exchange_equals_barrier_exchange)


def test_train_loop_torch_engine_symmetric_bucket_and_flat_sgd():
    go(W.w_train_torch_engine_gpu)


def test_bf16_wire_exchange_tracks_fp32_wire():
    go(W.w_bf16_wire_exchange)


def test_batched_tensor_core_trainer_equals_global_batch_sgd():
    go(W.w_batched_trainer)


def test_subgroup_symmetric_world():
    go(W.w_subgroup_symmetric_world, size=max(2, min(_n(), 4)))


@pytest.mark.parametrize("world", [3, 5, 6, 7])
def test_odd_worlds(world):
    """gloo.py:59 runs 7 ranks; two-shot / NVLS slicing and the inbox layout must not assume a power of two."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    go(W.w_suite_odd_world, size=world)


def test_flag_reuse_stress_1e5():
    go(W.w_flag_reuse_stress)


def test_large_messages_vs_nccl_to_1gib():
    go(W.w_large_sizes_vs_nccl)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="the one-launch suite is for the 8-GPU box")
def test_world8_suite_one_launch():
    go(W.w_suite_world, size=8)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="the one-launch suite is for the 8-GPU box")
def test_world8_suite_rest_one_launch():
    go(W.w_suite_world_rest, size=8)
