"""Step results of a sharded batch -> the agent's rank ("one NCCL gather of step results").

One process per GPU steps its own block of instances; what the agent's rank needs from the others every step is small
(``rho``: 4 bytes per line and instance).  Issuing one collective per 45-microsecond step from Python costs more than the
step itself, so the results go into a device ring of ``2 * K`` step slots and ONE gather per ``K`` steps ships a half ring,
asynchronously, while the kernels fill the other half.  Backend-agnostic (``torch.distributed``: NCCL on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

__all__ = ["RingCollector"]


class RingCollector:
    def __init__(self, rank: int, world: int, steps_per_gather: int, slot_shape, device, dtype=torch.float32, dst: int = 0):
        self.rank, self.world, self.K, self.dst = int(rank), int(world), max(1, int(steps_per_gather)), int(dst)
        self.ring = torch.zeros((2 * self.K,) + tuple(slot_shape), dtype=dtype, device=device)
        self.lists: Optional[List[List[torch.Tensor]]] = None
        if self.world > 1 and self.rank == self.dst:
            self.lists = [[torch.empty((self.K,) + tuple(slot_shape), dtype=dtype, device=device) for _ in range(self.world)]
                          for _ in range(2)]
        self.works = []
        self.gathered_upto = 0
        self.n_steps = 0

    def slot(self, k: int) -> torch.Tensor:
        """where step k's results go (the kernels write here)"""
        return self.ring[k % (2 * self.K)]

    def step_done(self, k: int) -> None:
        """call right after step k has been enqueued; launches the gather of a half ring when step k completes it"""
        self.n_steps = k + 1
        if self.world <= 1 or (k + 1) % self.K:
            return
        half = (k % (2 * self.K)) // self.K
        self.works.append(dist.gather(self.ring[half * self.K:(half + 1) * self.K], self.lists[half] if self.lists else None,
                                      dst=self.dst, async_op=True))
        while len(self.works) > 1:          # the next steps refill the other half: its gather (K steps old) must be through
            self.works.pop(0).wait()
        self.gathered_upto = k + 1

    def drain(self) -> None:
        """wait for what is in flight and gather the partly filled half"""
        while self.works:
            self.works.pop(0).wait()
        if self.world > 1 and self.gathered_upto < self.n_steps:
            half = ((self.n_steps - 1) % (2 * self.K)) // self.K
            dist.gather(self.ring[half * self.K:(half + 1) * self.K], self.lists[half] if self.lists else None, dst=self.dst)
            self.gathered_upto = self.n_steps

    def gathered(self, r: int, k: int) -> torch.Tensor:
        """on the agent's rank, after :meth:`drain`: rank r's results of step k (k within the last K..2K steps)"""
        if self.world <= 1:
            return self.ring[k % (2 * self.K)]
        s = k % (2 * self.K)
        return self.lists[s // self.K][r][s % self.K]
