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
