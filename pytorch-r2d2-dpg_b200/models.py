"""Drop-in for the reference's models.py: ActorNet / CriticNet with the same constructor, stateful
per-step `__call__`, set_state / reset_state / get_state and state_dict keys (models.py:14-97).

These modules serve the ACTOR side (batch-1 inference, out of the hot-path scope) and checkpoint
exchange; the learner never steps them - it runs whole windows through the CUDA chains
(r2d2_b200.engine.LearnerEngine) on flat parameter blocks that share this state_dict layout.
The hidden size is a keyword (default 128 = the reference's literal).
"""
import numpy as np
import torch
import torch.nn as nn

HIDDEN_DEFAULT = 128


def _fanin_uniform(t):
    bound = 1.0 / np.sqrt(t.size(0))  # keyed on out_features, as the reference does (models.py:8-11)
    return torch.empty_like(t).uniform_(-bound, bound)


class _RecurrentNet(nn.Module):
    _critic = False

    def __init__(self, obs_size, n_actions, cuda_id=0, hidden=HIDDEN_DEFAULT):
        super().__init__()
        self.hidden = hidden
        self.l1 = nn.Linear(obs_size + (n_actions if self._critic else 0), hidden)
        self.l2 = nn.LSTMCell(hidden, hidden)
        self.l3 = nn.Linear(hidden, n_actions)
        with torch.no_grad():
            self.l1.weight.copy_(_fanin_uniform(self.l1.weight))
            self.l2.weight_ih.copy_(_fanin_uniform(self.l2.weight_ih))
            self.l2.weight_hh.copy_(_fanin_uniform(self.l2.weight_hh))
            self.l3.weight.uniform_(-3e-3, 3e-3)
            self.l3.bias.fill_(3e-4)
        self.hx = self.cx = None
        self.cuda_id = cuda_id

    def _cell(self, x):
        z = torch.tanh(self.l1(x))
        if self.hx is None:
            self.hx = z.new_zeros((z.size(0), self.hidden))
            self.cx = z.new_zeros((z.size(0), self.hidden))
        self.hx, self.cx = self.l2(z, (self.hx, self.cx))
        return self.hx

    def set_state(self, hx, cx):
        self.hx, self.cx = hx, cx

    def reset_state(self):
        self.hx = self.cx = None

    def get_state(self):
        if self.hx is None:
            return (np.zeros((1, self.hidden), dtype=np.float32), np.zeros((1, self.hidden), dtype=np.float32))
        return self.hx.detach().cpu().numpy().copy(), self.cx.detach().cpu().numpy().copy()


class ActorNet(_RecurrentNet):
    def __call__(self, x):
        return torch.tanh(self.l3(torch.tanh(self._cell(x))))


class CriticNet(_RecurrentNet):
    _critic = True

    def __call__(self, x, a):
        return self.l3(self._cell(torch.cat((x, a), 1)))  # head on hx itself, n_actions outputs (models.py:61,82)
