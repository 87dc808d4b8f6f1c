"""One learner process per GPU: process-group bootstrap, actor-file sharding and gradient averaging.

The reference spawns ONE learner (r2d2.py:19) that polls every actor's file (learner.py:69-75,144-149).  The
data-parallel form of the path is: rank r of W owns GPU LOCAL_RANK, the replay shard fed by the actors
{i : i mod W == r}, and averages the two flat gradient blocks with the other ranks at the two optimiser steps
(learner.py:114,128) - nothing else crosses GPUs.  This module is pure host logic (no CUDA calls at import) so that
the partition and the reduction can be tested on CPU with the gloo backend.
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass(frozen=True)
class DistEnv:
    rank: int = 0
    world: int = 1
    local_rank: int = 0

    @staticmethod
    def from_environ(env=None) -> "DistEnv":
        """torchrun / torch.distributed.run convention: RANK, WORLD_SIZE, LOCAL_RANK (absent -> single process)."""
        env = os.environ if env is None else env
        world = int(env.get("WORLD_SIZE", "1"))
        rank = int(env.get("RANK", "0"))
        local = int(env.get("LOCAL_RANK", str(rank)))
        if not (0 <= rank < world):
            raise ValueError(f"RANK={rank} outside WORLD_SIZE={world}")
        return DistEnv(rank=rank, world=world, local_rank=local)

    @property
    def distributed(self) -> bool:
        return self.world > 1

    @property
    def is_main(self) -> bool:
        return self.rank == 0

    def owned_actors(self, n_actors: int) -> list:
        """Actor ids whose memory{i}.pt this rank ingests: i mod world == rank (disjoint, covers range(n_actors))."""
        return list(range(self.rank, n_actors, self.world))

    def init_process_group(self, backend: str = "nccl", device=None):
        """Join the job's process group (idempotent).  Rendezvous defaults to 127.0.0.1 (single node)."""
        import torch.distributed as dist
        if not self.distributed:
            return None
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)
        if dist.get_world_size() != self.world or dist.get_rank() != self.rank:
            raise RuntimeError("process group does not match RANK / WORLD_SIZE")
        return dist


class GradSync:
    """Average flat gradient blocks over the ranks.  On CUDA the all-reduce runs on a side stream so that the caller
    can keep the compute stream busy with work that does not read the gradients (the actor's forward chain while the
    critic gradients are reduced); `wait()` orders the compute stream behind it.  The SUM is reduced; the division by
    the world size is folded into the optimiser kernel (grad_scale)."""

    def __init__(self, dist, world: int):
        self.dist, self.world = dist, world
        self._stream = None
        self._done = None

    def start(self, flat):
        import torch
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            cur = torch.cuda.current_stream(flat.device)
            self._stream.wait_stream(cur)                    # the gradients are complete on the compute stream
            with torch.cuda.stream(self._stream):
                self.dist.all_reduce(flat)
                self._done = torch.cuda.Event()
                self._done.record(self._stream)
            flat.record_stream(self._stream)
        else:
            self.dist.all_reduce(flat)

    def wait(self, device=None):
        import torch
        if self._done is not None:
            torch.cuda.current_stream(device).wait_event(self._done)
            self._done = None

    def replicas_identical(self, tensors) -> bool:
        """True when every rank holds bit-identical copies of the given tensors (data-parallel invariant)."""
        import torch
        ok = True
        for t in tensors:
            bits = t.detach().contiguous().view(torch.int32)
            hi, lo = bits.clone(), bits.clone()
            self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
            ok = ok and bool(torch.equal(hi, lo))
        return ok
