"""Tiny NCCL all-reduce under NCCL_DEBUG=INFO (shows whether NVLS is in use on this box)."""
import os
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(1 << 24, device="cuda")
dist.all_reduce(t)
torch.cuda.synchronize()
dist.destroy_process_group()
