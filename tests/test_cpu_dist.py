"""Host logic of the data-parallel learner on CPU with the gloo backend, world size 2: the environment
bootstrap the drop-in learner uses, the actor-file partition (rank r ingests actors i = r mod W; the reference's single
learner polls all of them, learner.py:69-75,144-149), the gradient average that replaces nothing in the reference
(one learner there) and must equal the global-batch mean, and the replica-identity check bench.py prints."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-r2d2-dpg_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def test_dist_env_parsing_and_partition():
    from r2d2_b200.dist_env import DistEnv
    assert DistEnv.from_environ({}) == DistEnv(0, 1, 0)
    e = DistEnv.from_environ({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"})
    assert e.distributed and not e.is_main and e.owned_actors(16) == [3, 11]
    for world in (1, 2, 3, 8):
        for n_actors in (1, 5, 16, 64):
            owned = [DistEnv(r, world, r).owned_actors(n_actors) for r in range(world)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(n_actors))                      # disjoint cover of every actor file
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    with pytest.raises(ValueError):
        DistEnv.from_environ({"RANK": "2", "WORLD_SIZE": "2"})


def _worker(rank, world, port, n_actors):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from r2d2_b200.dist_env import DistEnv, GradSync
    env = DistEnv.from_environ()
    dist = env.init_process_group("gloo")
    assert dist.get_rank() == rank and env.owned_actors(n_actors) == list(range(rank, n_actors, world))
    # every actor file is owned exactly once across the job
    mine = torch.zeros(n_actors, dtype=torch.int32)
    mine[env.owned_actors(n_actors)] = 1
    dist.all_reduce(mine)
    assert bool((mine == 1).all())
    # gradient average: per-rank gradients of equal batch shards -> mean == global-batch gradient
    g = torch.Generator().manual_seed(7)
    per_rank = [torch.randn(1000, generator=g) for _ in range(world)]
    flat = per_rank[rank].clone()
    sync = GradSync(dist, world)
    sync.start(flat)
    sync.wait()
    want = torch.stack(per_rank).sum(0)
    assert torch.allclose(flat, want, rtol=0, atol=1e-6)              # SUM is reduced; 1/world is the optimiser's grad_scale
    assert torch.allclose(flat / world, torch.stack(per_rank).mean(0), atol=1e-6)
    # replica identity check: identical tensors pass, one differing bit fails
    same = torch.arange(64, dtype=torch.float32)
    assert sync.replicas_identical([same, same * 2])
    diff = same.clone()
    if rank == 1:
        diff[5] = torch.nextafter(diff[5], torch.tensor(1e9))
    assert not sync.replicas_identical([diff])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partition_and_gradient_average():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 200, 5), nprocs=2, join=True)
