"""Hierarchical all-reduce for jobs that span several machines.

The peer-memory kernels (``parallel/symm.py``) live inside ONE NVSwitch domain: every GPU of the machine maps every other
GPU's memory.  Across machines there is no such mapping, so a multi-node ``backend="b200"`` job gets a two-level world:

  level 1  inside each machine: the fused peer-memory all-reduce over the machine's ranks (``SymmWorld`` over a sub-group;
           LL / one-shot / two-shot / NVLS by size, the ``scale`` fused), exactly the single-node path;
  level 2  across machines: one NCCL all-reduce per *rail* -- the ranks with the same local index on every machine form a
           group, and rail ``l`` reduces slice ``l`` of the buffer (1/L of the bytes per rank over the network, all rails in
           parallel), followed by a level-1 all-gather of the slices (NCCL inside the machine).

``HierWorld`` offers the part of ``SymmWorld``'s interface the gradient path uses (``supports``, ``alloc``, ``all_reduce_``,
``describe``, ``destroy``), so ``average_gradients``, ``GradBucket`` / ``DistributedDataParallel`` and ``comm.all_reduce``
work unchanged.  The in-kernel exchange of the fused ConvNet trainers is a single-domain protocol; on several machines
``train()`` uses the bucketed engine instead.

The reference is a single-machine program (train_dist.py:138-147 forks its ranks on localhost); its tutorial's init-method
section (tuto.md:404-428) is where several machines enter.  Tested on the CPU tier with gloo and simulated host names
(``B200DIST_FAKE_HOSTNAME``); the GPU tiers of this repo run on one machine.
"""
from __future__ import annotations

import os
import socket
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

__all__ = ["HierWorld", "hostname", "node_layout", "init_hier_world"]


def hostname() -> str:
    """This rank's machine name (``B200DIST_FAKE_HOSTNAME`` overrides it: tests simulate several machines on one)."""
    return os.environ.get("B200DIST_FAKE_HOSTNAME") or socket.gethostname()


def node_layout(hosts: List[str]) -> Dict[str, object]:
    """From the gathered host names (index = global rank): machines in first-appearance order, the ranks of each, and
    whether the layout is regular (same number of ranks on every machine -- needed for the per-rail level 2)."""
    order: List[str] = []
    by_host: Dict[str, List[int]] = {}
    for r, h in enumerate(hosts):
        if h not in by_host:
            by_host[h] = []
            order.append(h)
        by_host[h].append(r)
    nodes = [by_host[h] for h in order]
    return {"hosts": order, "nodes": nodes, "regular": len({len(n) for n in nodes}) == 1}


class HierWorld:
    """Two-level all-reduce world over the default process group (see the module docstring)."""

    def __init__(self, hosts: Optional[List[str]] = None):
        if hosts is None:
            hosts = [None] * dist.get_world_size()
            dist.all_gather_object(hosts, hostname())
        lay = node_layout(hosts)
        if not lay["regular"]:
            raise RuntimeError(f"hierarchical world needs the same number of ranks on every machine, got {lay['nodes']}")
        self.nodes: List[List[int]] = lay["nodes"]
        self.hosts: List[str] = lay["hosts"]
        self.global_rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.n_nodes, self.local_size = len(self.nodes), len(self.nodes[0])
        self.node = next(i for i, n in enumerate(self.nodes) if self.global_rank in n)
        self.local_rank = self.nodes[self.node].index(self.global_rank)
        # every rank creates every group, in the same order (torch.distributed's rule)
        self.local_group = None
        for i, n in enumerate(self.nodes):
            g = dist.new_group(n)
            if i == self.node:
                self.local_group = g
        self.rail_group = None
        for l in range(self.local_size):
            g = dist.new_group([n[l] for n in self.nodes])
            if l == self.local_rank:
                self.rail_group = g
        self.local: Optional[object] = None            # SymmWorld over the machine's ranks (CUDA only)
        self.device = torch.device("cpu")
        if torch.cuda.is_available() and "nccl" in str(dist.get_backend()):
            from . import symm
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.local = symm.init_world(self.local_group)
        # interface parity with SymmWorld
        self.world = self.size
        self.multicast = bool(self.local is not None and self.local.multicast)

    # ------------------------------------------------------------------ SymmWorld-compatible surface
    def supports(self, t: torch.Tensor) -> bool:
        if self.local is not None:
            return self.local.supports(t)
        return (not t.is_cuda) and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16, torch.float64)

    def alloc(self, numel: int, dtype: torch.dtype):
        """A gradient bucket: symmetric inside the machine when the peer-memory world exists, plain memory otherwise."""
        if self.local is not None:
            return self.local.alloc(numel, dtype)
        n = (numel + 63) // 64 * 64

        class _Plain:
            pass
        hd = _Plain()
        hd.local, hd.numel, hd.dtype = torch.zeros(n, dtype=dtype), n, dtype
        return hd

    def all_reduce_(self, t: torch.Tensor, scale: float = 1.0, handle=None, variant: Optional[int] = None,
                    wire: Optional[torch.dtype] = None, max_blocks: Optional[int] = None) -> torch.Tensor:
        """In-place ``t <- scale * sum over ALL ranks of t`` (level 1 fused kernel, level 2 per-rail NCCL, level-1 gather)."""
        if not self.supports(t):
            raise TypeError("hierarchical all_reduce needs a contiguous float tensor on this rank's device")
        flat = t.view(-1)
        # level 1: inside the machine (scale fused)
        if self.local is not None:
            self.local.all_reduce_(flat, scale=scale, handle=handle, variant=variant, wire=wire, max_blocks=max_blocks)
        else:
            if self.local_size > 1:
                dist.all_reduce(flat, group=self.local_group)
            if scale != 1.0:
                flat.mul_(scale)
        if self.n_nodes == 1:
            return t
        # level 2: rail l reduces slice l across the machines, then the slices are gathered inside the machine
        L, n = self.local_size, flat.numel()
        per = (n + L - 1) // L
        if L == 1 or n < 4096:                           # tiny message: one collective over the rail beats three
            dist.all_reduce(flat, group=self.rail_group)
            return t
        lo, hi = min(n, self.local_rank * per), min(n, (self.local_rank + 1) * per)
        if hi > lo:
            dist.all_reduce(flat[lo:hi], group=self.rail_group)
        if n == per * L:
            dist.all_gather_into_tensor(flat, flat[lo:hi].clone(), group=self.local_group)
        else:                                            # ragged tail: broadcast every slice from its owner
            for l in range(L):
                a, b = min(n, l * per), min(n, (l + 1) * per)
                if b > a:
                    dist.broadcast(flat[a:b], src=self.nodes[self.node][l], group=self.local_group)
        return t

    def describe(self) -> dict:
        d = {"hierarchical": True, "world": self.size, "nodes": self.n_nodes, "ranks_per_node": self.local_size,
             "node": self.node, "local_rank": self.local_rank, "hosts": self.hosts}
        if self.local is not None:
            d["local"] = self.local.describe()
        return d

    def destroy(self):
        self.local = None                                # the local SymmWorld is registered (and destroyed) on its own


def init_hier_world() -> HierWorld:
    """Collective over the default group: build the two-level world and register it as THE world of the default group, so
    that ``comm.all_reduce`` / ``average_gradients`` / ``GradBucket`` find it where they look for a ``SymmWorld``."""
    from . import symm
    w = symm._WORLDS.get(None)
    if isinstance(w, HierWorld):
        return w
    w = HierWorld()
    symm._WORLDS[None] = w
    return w
