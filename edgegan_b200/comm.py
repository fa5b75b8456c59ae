"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink) all-reduce of the
flat gradient buffers and of the generator's sync-BN sums."""
from __future__ import annotations

import os


class TorchDistComm:
    def __init__(self, backend=None):
        import torch.distributed as dist
        self.dist = dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend or "nccl")
        self.world_size = dist.get_world_size()
        self.rank = dist.get_rank()

    def allreduce(self, t):
        self.dist.all_reduce(t)          # sum
        return t

    def barrier(self):
        self.dist.barrier()
