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


def _worker(rank, world, port, out_dir, mode):
    import torch.distributed as dist
    from r2d2_b200 import engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), R2D2_DP_MODE=mode)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    B = 16
    pc = ref_port.PathConfig(obs=6, act=2, hidden=64, batch=B, burn_in=4, learning=6, n_step=2)
    cfg = engine.PathConfig(obs=6, act=2, hidden=64, batch=B // world, burn_in=4, learning=6, n_step=2)
    eng = engine.LearnerEngine(cfg, device=f"cuda:{rank}", seed=5)
    eng.enable_data_parallel()
    assert eng._dp_mode == mode, f"gradient exchange fell back to {eng._dp_mode}"
    init = {n: {k: v.cpu().numpy().copy() for k, v in eng.views(n).items()} for n in ("actor", "critic")}
    sh = B // world
    for it in range(3):
        full = ref_port.synthetic_batch(pc, seed=10 + it)
        shard = {k: np.ascontiguousarray(v[:, rank * sh:(rank + 1) * sh]) for k, v in full.items()}
        eng.set_batch(shard)
        eng.step()
    torch.cuda.synchronize()
    assert eng.replicas_identical() and eng.peer_status() == 0
    np.savez(os.path.join(out_dir, f"dp_rank{rank}.npz"), **{f"{n}/{k}": v.cpu().numpy() for n in ("actor", "critic")
                                                            for k, v in eng.views(n).items()})
    if rank == 0:
        cfg1 = engine.PathConfig(obs=6, act=2, hidden=64, batch=B, burn_in=4, learning=6, n_step=2)
        single = engine.LearnerEngine(cfg1, device="cuda:0", seed=5)
        single.load_state_dicts(init["actor"], init["critic"])
        for it in range(3):
            single.set_batch(ref_port.synthetic_batch(pc, seed=10 + it))
            single.step()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "single.npz"), **{f"{n}/{k}": v.cpu().numpy() for n in ("actor", "critic")
                                                         for k, v in single.views(n).items()},
                 **{f"init/{n}/{k}": v for n in init for k, v in init[n].items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "defer"])
def test_two_gpu_data_parallel_matches_single_gpu(mode):
    """mode "peer": the library's own signal / slice-sum / wait kernels over NVLink peer memory (csrc/peer.cu, the
    default); "defer": NCCL all-reduces on a side stream (the fallback)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, 29600 + os.getpid() % 200 + (7 if mode == "peer" else 0), d, mode), nprocs=2, join=True)
        r0, r1, one = (np.load(os.path.join(d, f)) for f in ("dp_rank0.npz", "dp_rank1.npz", "single.npz"))
        for k in r0.files:
            assert np.array_equal(r0[k], r1[k]), f"replicas diverged: {k}"
            upd, ref = r0[k] - one["init/" + k], one[k] - one["init/" + k]
            assert rel_l2(r0[k], one[k]) < 1e-4, k                      # parameters
            assert rel_l2(upd, ref) < 5e-2, k                           # and the (Adam, sign-like) updates agree


def test_torchrun_dropin_learner():
    """The PRODUCT path, not a harness: `learner.Learner` launched once per GPU by torch.distributed.run binds
    cuda:LOCAL_RANK, joins NCCL, ingests only the actor files i = rank mod world (the reference's single learner polls all
    of them, learner.py:69-75,144-149), keeps its replicas bit-identical and lets rank 0 alone write model.pt."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "pytorch-r2d2-dpg_b200")]
    env = dict(os.environ, R2D2_OBS_SIZE="5", R2D2_N_ACTIONS="2", R2D2_HIDDEN="64", R2D2_BATCH="4",
               R2D2_ACTOR_DEVICE="cpu")
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "model_data"))
        os.makedirs(os.path.join(d, "memory_data"))
        # four CPU actors write episode files in the reference's format (actor.py:163-179)
        code = ("import sys; sys.path[:0]=[%r,%r]; import actor\n"
                "for aid in range(4):\n"
                "    a = actor.Actor(aid); a.env.episode_len = 150; a.run(max_episodes=5)\n") % (
                    root, os.path.join(root, "pytorch-r2d2-dpg_b200"))
        # actors need a model.pt to follow: a single-process learner writes the initial one
        init = ("import sys; sys.path[:0]=[%r,%r]; import learner; learner.Learner(4)\n") % (
            root, os.path.join(root, "pytorch-r2d2-dpg_b200"))
        subprocess.check_call([sys.executable, "-c", init], cwd=d, env=env)
        subprocess.check_call([sys.executable, "-c", code], cwd=d, env=env)
        os.remove(os.path.join(d, "model_data", "model.pt"))
        port = 29800 + os.getpid() % 100
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                               "--master-addr", "127.0.0.1", "--master-port", str(port),
                               os.path.join(root, "tests", "dp_dropin_worker.py"), "4", "4"], cwd=d, env=env, timeout=600)
        r = [json.load(open(os.path.join(d, "dp_rank%d.json" % i))) for i in range(2)]
        assert [x["device"] for x in r] == ["cuda:0", "cuda:1"]
        assert r[0]["owned"] == [0, 2] and r[1]["owned"] == [1, 3]
        assert r[0]["episodes"] > 0 and r[1]["episodes"] > 0
        assert r[0]["steps"] == r[1]["steps"] == 4
        assert r[0]["replicas_identical"] and r[1]["replicas_identical"]
        assert r[0]["param_sum"] == r[1]["param_sum"]
        assert os.path.isfile(os.path.join(d, "model_data", "model.pt"))        # written by rank 0 alone
