"""Reads off the TMEM lane <-> accumulator-row mapping of a UMMA_M=64 tcgen05.mma (cta_group::1) on this GPU."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_tuto.pth_b200.ops import _ext  # noqa: E402

C = _ext.C()
dev = torch.device("cuda", 0)
a = torch.zeros(64, 64, dtype=torch.bfloat16, device=dev)
a[:, 0] = torch.arange(1, 65, dtype=torch.float32, device=dev).to(torch.bfloat16)
b = torch.zeros(32, 64, dtype=torch.bfloat16, device=dev)
b[:, 0] = 1.0
b[:, 1] = torch.arange(32, device=dev).to(torch.bfloat16)
a[:, 1] = 0.0
dump = C.gemm_probe_m64(a, b)
torch.cuda.synchronize()
rows = dump[:, 0].round().to(torch.int64).tolist()        # lane -> (row + 1) or 0
mapping = {lane: r - 1 for lane, r in enumerate(rows) if r > 0}
cols_ok = bool(torch.all(dump[list(mapping.keys())][:, 1:] == dump[list(mapping.keys())][:, :1]))
print(json.dumps({"lanes_used": sorted(mapping.keys()), "lane_to_row": mapping, "all_columns_equal": cols_ok}))
