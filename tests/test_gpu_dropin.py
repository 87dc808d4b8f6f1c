"""The drop-in module surface on the GPU: actor files in the reference's format -> learner.Learner (same constructor,
attributes and file protocol as learner.py:22-67) ingests them, runs iterations, writes model.pt that actors reload."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_learner_dropin_end_to_end(monkeypatch):
    monkeypatch.setenv("R2D2_OBS_SIZE", "5")
    monkeypatch.setenv("R2D2_N_ACTIONS", "2")
    monkeypatch.setenv("R2D2_HIDDEN", "64")
    monkeypatch.setenv("R2D2_BATCH", "4")
    for m in ("actor", "learner", "replay_memory", "models", "utils"):
        sys.modules.pop(m, None)
    import actor as dropin_actor
    import learner as dropin_learner
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        try:
            os.makedirs("model_data")
            os.makedirs("memory_data")
            lr = dropin_learner.Learner(n_actors=2)                       # writes model.pt (learner.py:54)
            assert os.path.isfile("model_data/model.pt")
            sd = torch.load("model_data/model.pt")
            assert set(sd) == {"actor", "target_actor", "critic", "target_critic"}
            assert list(sd["actor"].keys()) == ["l1.weight", "l1.bias", "l2.weight_ih", "l2.weight_hh", "l2.bias_ih",
                                                "l2.bias_hh", "l3.weight", "l3.bias"]
            for aid in range(2):                                           # two CPU actors follow model.pt and write episodes
                a = dropin_actor.Actor(aid)
                a.env.episode_len = 150
                a.run(max_episodes=5)
                assert os.path.isfile(f"memory_data/memory{aid}.pt")
            lr.model_save_interval = 2
            lr.memory_update_interval = 2
            lr.run(max_steps=4)                                            # gate: >= 100*batch sequences (learner.py:69)
            assert lr.memory.sequence_counter >= 400
            assert lr.engine.step_count == 4
            assert len(lr.memory.priority) == len(lr.memory.memory) >= 8
            p00 = lr.memory.priority[0][0]
            assert np.isfinite(p00) and p00 >= 0
            assert abs(lr.memory.total_priority[0] - sum(lr.memory.priority[0])) < 1e-3
            out = lr.memory.sample()                                       # the reference's 10-tuple (replay_memory.py:135-136)
            assert len(out) == 10 and out[2].shape == (65, 4, 5) and out[6].shape == (2, 4, 64) and out[2].is_cuda
            e, s = out[0][0], out[1][0]
            assert 0 <= e < len(lr.memory.memory) and 0 <= s < len(lr.memory.priority[e])
            a = dropin_actor.Actor(0)                                      # reloads the learner's checkpoint
            for k, v in lr.engine.views("actor").items():
                assert torch.allclose(a.actor.state_dict()[k], torch.load("model_data/model.pt")["actor"][k].cpu())
        finally:
            os.chdir(cwd)
            for m in ("actor", "learner", "replay_memory", "models", "utils"):
                sys.modules.pop(m, None)
