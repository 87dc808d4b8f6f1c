"""Worker of tests/test_gpu_multigpu.py::test_torchrun_dropin_learner: one process per GPU under torch.distributed.run.
Builds the drop-in `learner.Learner` exactly as `learner_process` does (learner.py:18-20), lets it ingest the actor
files it owns, runs a few iterations and writes what the test checks into <cwd>/dp_rank{r}.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-r2d2-dpg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import learner as dropin_learner  # noqa: E402


def main():
    n_actors, steps = int(sys.argv[1]), int(sys.argv[2])
    lr = dropin_learner.Learner(n_actors)
    lr.model_save_interval = 2
    lr.memory_update_interval = 1000
    lr.run(max_steps=steps)
    torch.cuda.synchronize()
    eng = lr.engine
    out = {"rank": lr.dist_env.rank, "world": lr.dist_env.world, "device": str(eng.device),
           "owned": lr.dist_env.owned_actors(n_actors), "episodes": len(lr.memory.memory),
           "sequence_counter": int(lr.memory.sequence_counter), "steps": eng.step_count,
           "replicas_identical": eng.replicas_identical(),
           "param_sum": float(eng.flat["actor"].double().sum().item() + eng.flat["critic"].double().sum().item()),
           "leaf0": int(eng.leaf_idx[0].item())}
    with open("dp_rank{}.json".format(lr.dist_env.rank), "w") as f:
        json.dump(out, f)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
