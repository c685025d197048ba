"""Data-parallel engine: gradient buckets, overlapped fused all-reduce, symmetric peer memory."""
from .ddp import GradBucket, average_gradients, DistributedDataParallel, broadcast_parameters  # noqa: F401
