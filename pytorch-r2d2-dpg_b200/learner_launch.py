"""One learner rank: `torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 learner_launch.py [n_actors]`
runs `learner.learner_process(n_actors)` (learner.py:18-20) once per GPU; with no torchrun it is the reference's single
learner process.  Actors are started separately (r2d2.py starts both for the single-learner case)."""
import sys

from learner import learner_process

if __name__ == "__main__":
    learner_process(int(sys.argv[1]) if len(sys.argv) > 1 else 16)   # r2d2.py:13 n_actors = 16
