"""Data-parallel learner on >= 2 GPUs (NCCL): equal batch shards + all-reduce of the two flat gradient blocks reproduce
the single-GPU update on the global batch.  Skipped on single-GPU boxes."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import ref_port

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from r2d2_b200 import engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    B = 16
    pc = ref_port.PathConfig(obs=6, act=2, hidden=64, batch=B, burn_in=4, learning=6, n_step=2)
    cfg = engine.PathConfig(obs=6, act=2, hidden=64, batch=B // world, burn_in=4, learning=6, n_step=2)
    eng = engine.LearnerEngine(cfg, device=f"cuda:{rank}", seed=5)
    eng.enable_data_parallel()
    init = {n: {k: v.cpu().numpy().copy() for k, v in eng.views(n).items()} for n in ("actor", "critic")}
    sh = B // world
    for it in range(2):
        full = ref_port.synthetic_batch(pc, seed=10 + it)
        shard = {k: np.ascontiguousarray(v[:, rank * sh:(rank + 1) * sh]) for k, v in full.items()}
        eng.set_batch(shard)
        eng.step()
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"dp_rank{rank}.npz"), **{f"{n}/{k}": v.cpu().numpy() for n in ("actor", "critic")
                                                            for k, v in eng.views(n).items()})
    if rank == 0:
        cfg1 = engine.PathConfig(obs=6, act=2, hidden=64, batch=B, burn_in=4, learning=6, n_step=2)
        single = engine.LearnerEngine(cfg1, device="cuda:0", seed=5)
        single.load_state_dicts(init["actor"], init["critic"])
        for it in range(2):
            single.set_batch(ref_port.synthetic_batch(pc, seed=10 + it))
            single.step()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "single.npz"), **{f"{n}/{k}": v.cpu().numpy() for n in ("actor", "critic")
                                                         for k, v in single.views(n).items()},
                 **{f"init/{n}/{k}": v for n in init for k, v in init[n].items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_data_parallel_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, 29600 + os.getpid() % 200, d), nprocs=2, join=True)
        r0, r1, one = (np.load(os.path.join(d, f)) for f in ("dp_rank0.npz", "dp_rank1.npz", "single.npz"))
        for k in r0.files:
            assert np.array_equal(r0[k], r1[k]), f"replicas diverged: {k}"
            upd, ref = r0[k] - one["init/" + k], one[k] - one["init/" + k]
            assert rel_l2(r0[k], one[k]) < 1e-4, k                      # parameters
            assert rel_l2(upd, ref) < 5e-2, k                           # and the (Adam, sign-like) updates agree
