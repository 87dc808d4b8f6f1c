"""Drop-in for the reference's replay_memory.py.

  ReplayMemory          actor-side episode buffer + `save(actorID)` in the reference's file format
                        (replay_memory.py:8-64); pure host code, unchanged contract.
  LearnerReplayMemory   same constructor / attributes / `sample()` 10-tuple / `load(actorID)` as
                        replay_memory.py:67-175, but the storage is a replay shard in HBM with a sum tree
                        (r2d2_b200.engine.DeviceReplay -> csrc/replay.cu).  P(episode, sequence) is
                        proportional to priority[episode][sequence], the law of the reference's two-level
                        WeightedRandomSampler draw (replay_memory.py:95-114).

File format (actor.py:163-176, replay_memory.py:55-59): torch.save of
  {'replay_memory': deque[list[(obs f32[O], act f32[A], [reward], [terminal])]],
   'recurrent_state': deque[list[[[hx, cx] x 4 nets]]], 'priority': deque[list[float]], 'total_priority': list}.
"""
import os
from collections import deque
from time import sleep

import numpy as np
import torch

BURN_IN, LEARNING, N_STEP = 20, 40, 5  # replay_memory.py:18-20,78-81


class ReplayMemory:
    def __init__(self, memory_sequence_size=100000, batch_size=64):
        self.path = './memory_data/'
        self.memory_sequence_size = memory_sequence_size
        self.batch_size = batch_size
        self.memory, self.priority = deque(), deque()
        self.total_priority, self.recurrent_state = deque(), deque()
        self.burn_in_length, self.learning_length, self.n_step = BURN_IN, LEARNING, N_STEP

    def add(self, episode, recurrent_state, priority):
        self.memory.append(episode)
        self.recurrent_state.append(recurrent_state)
        self.priority.append(priority)
        self.total_priority.append(sum(list(priority)))

    def clear(self):
        for d in (self.memory, self.recurrent_state, self.priority, self.total_priority):
            d.clear()

    def size(self):
        return sum(len(ep) for ep in self.memory)

    def save(self, actorID):
        """Write (the reference always overwrites: its append branch tests a file that never exists,
        replay_memory.py:38) and clear.  Written to a temp name and renamed so a concurrently loading
        learner never sees a half-written pickle."""
        payload = {'replay_memory': self.memory, 'recurrent_state': self.recurrent_state,
                   'priority': self.priority, 'total_priority': list(self.total_priority)}
        final = self.path + 'memory{}.pt'.format(actorID)
        tmp = final + '.tmp{}'.format(os.getpid())
        torch.save(payload, tmp)
        os.replace(tmp, final)
        self.clear()


def pack_episode(rows, states, hidden=None):
    """One episode of the actor format -> contiguous float32 arrays (obs, act, rew, term, states[E,4,2,H])."""
    obs = np.stack([np.asarray(r[0], np.float32) for r in rows])
    act = np.stack([np.asarray(r[1], np.float32) for r in rows])
    rew = np.asarray([r[2][0] for r in rows], np.float32)
    term = np.asarray([r[3][0] for r in rows], np.float32)
    st = np.asarray(states, np.float32)  # [E, 4, 2, H]
    if st.ndim != 4 or st.shape[1:3] != (4, 2):
        raise ValueError("recurrent_state must be [steps][4 nets][hx, cx][H], got %r" % (st.shape,))
    if hidden is not None and st.shape[3] != hidden:
        raise ValueError("recurrent state width %d != hidden %d" % (st.shape[3], hidden))
    return obs, act, rew, term, st


class _EpisodePriorities:
    """`memory.priority[e]`: indexable / assignable view of one episode's leaf priorities in HBM."""

    def __init__(self, owner, e):
        self._o, self._e = owner, e

    def __len__(self):
        return self._o._episodes[self._e][2]

    def _row(self, s):
        start, _, n_starts = self._o._episodes[self._e]
        if s < 0:
            s += n_starts
        if not 0 <= s < n_starts:
            raise IndexError(s)
        return start + s

    def __getitem__(self, s):
        return float(self._o._dev.tree_level(0)[self._row(s)].item())

    def __setitem__(self, s, value):
        o = self._o
        leaf = torch.tensor([self._row(s)], dtype=torch.int64, device=o._dev.device)
        o._dev.update_priorities(leaf, torch.tensor([float(value)], dtype=torch.float32, device=o._dev.device))

    def __iter__(self):
        start, _, n_starts = self._o._episodes[self._e]
        return iter(self._o._dev.tree_level(0)[start:start + n_starts].cpu().tolist())


class _PriorityTable:
    def __init__(self, owner):
        self._o = owner

    def __len__(self):
        return len(self._o._episodes)

    def __getitem__(self, e):
        return _EpisodePriorities(self._o, e)


class _TotalPriority:
    """`memory.total_priority[e]`: the tree maintains the sums, so assignment is accepted and ignored
    (learner.py:139 recomputes it from priority[e] after every write)."""

    def __init__(self, owner):
        self._o = owner

    def __len__(self):
        return len(self._o._episodes)

    def __getitem__(self, e):
        return float(sum(_EpisodePriorities(self._o, e)))

    def __setitem__(self, e, value):
        pass


class LearnerReplayMemory:
    def __init__(self, memory_sequence_size=500000, batch_size=32, obs_size=None, n_actions=None, hidden=128,
                 capacity_rows=None, device=None):
        self.path = './memory_data/'
        self.memory_sequence_size = memory_sequence_size
        self.sequence_counter = 0
        self.batch_size = batch_size
        self.burn_in_length, self.learning_length, self.n_step = BURN_IN, LEARNING, N_STEP
        self.sequence_length = self.burn_in_length + self.learning_length
        self._hidden, self._obs, self._act = hidden, obs_size, n_actions
        self._capacity_rows, self._device = capacity_rows, device
        self._dev = None          # DeviceReplay, created when the row width is known
        self._episodes = deque()  # (row_start, n_rows, n_starts) in FIFO order, mirrors the native ring
        self.priority = _PriorityTable(self)
        self.total_priority = _TotalPriority(self)

    # the reference exposes the raw deques; here the rows live in HBM
    @property
    def memory(self):
        return self._episodes

    @property
    def recurrent_state(self):
        return self._episodes

    def size(self):
        return sum(e[1] for e in self._episodes)

    def clear(self):
        if self._dev is not None:
            self._dev.close()
        self._dev = None
        self._episodes.clear()
        self.sequence_counter = 0

    def _cfg(self):
        from r2d2_b200.engine import PathConfig
        return PathConfig(obs=self._obs, act=self._act, hidden=self._hidden, batch=self.batch_size,
                          burn_in=self.burn_in_length, learning=self.learning_length, n_step=self.n_step)

    def _ensure_device(self, obs_size, n_actions, hidden):
        if self._dev is None:
            from r2d2_b200.engine import DeviceReplay
            self._obs, self._act, self._hidden = obs_size, n_actions, hidden
            rows = self._capacity_rows or int(min(self.memory_sequence_size, 2_000_000) * 1.3) + 4096
            self._dev = DeviceReplay(self._cfg(), capacity_rows=rows, max_sequences=self.memory_sequence_size,
                                     device=self._device)
        return self._dev

    def add_episode(self, rows, states, priority):
        obs, act, rew, term, st = pack_episode(rows, states)
        dev = self._ensure_device(obs.shape[1], act.shape[1], st.shape[3])
        dev.add_episode(obs, act, rew, term, st, np.asarray(priority, np.float32))
        stats = dev.stats()
        while len(self._episodes) + 1 > stats["n_episodes"]:   # native FIFO eviction happened
            self._episodes.popleft()
        self._episodes.append((int(stats["last_row_start"]), obs.shape[0], len(priority)))
        self.sequence_counter = int(stats["sequence_counter"])

    def get_weighted_sample_index(self):
        """Iterator of `batch_size` episode indices drawn proportionally to the episode totals
        (replay_memory.py:95-97), realised as the episode component of flat tree draws."""
        u = torch.rand(self.batch_size, device=self._dev.device)
        ep, _ = self._dev.decode(self._dev.sample_indices(u))
        return iter(int(e) for e in ep)

    def sample(self):
        """Same 10-tuple as replay_memory.py:99-136 (lists of ints + time-major CUDA tensors)."""
        B, T = self.batch_size, self.sequence_length + self.n_step
        dev = self._dev
        d = dev.device
        u = torch.rand(B, device=d)
        leaf = torch.empty(B, dtype=torch.int64, device=d)
        obs = torch.empty((T, B, self._obs), device=d)
        act = torch.empty((T, B, self._act), device=d)
        rew = torch.empty((T, B, 1), device=d)
        term = torch.empty((T, B, 1), device=d)
        states = torch.empty((4, 2, B, self._hidden), device=d)
        from r2d2_b200 import native as nv
        nv.check(dev.lib.r2d2_replay_sample(dev._h, nv.dptr(u), B, nv.dptr(leaf, torch.int64), nv.dptr(obs),
                                            nv.dptr(act), nv.dptr(rew), nv.dptr(term), nv.dptr(states),
                                            nv.current_stream()))
        ep, seq = dev.decode(leaf)
        return ([int(e) for e in ep], [int(s) for s in seq], obs, act, rew, term,
                states[0], states[1], states[2], states[3])

    def load(self, actorID):
        """Ingest memory{actorID}.pt and hand the file back emptied (replay_memory.py:138-157); one retry after
        a pause like the reference (replay_memory.py:158-175).  weights_only=False: the payload is a pickle of
        deques / ndarrays, which torch >= 2.6 refuses by default."""
        fname = self.path + 'memory{}.pt'.format(actorID)
        if not os.path.isfile(fname):
            return
        for attempt in (0, 1):
            try:
                payload = torch.load(fname, weights_only=False)
                for rows, states, prio in zip(payload['replay_memory'], payload['recurrent_state'], payload['priority']):
                    if len(rows) >= self.sequence_length + self.n_step:
                        self.add_episode(rows, states, prio)
                for key in ('replay_memory', 'recurrent_state', 'priority', 'total_priority'):
                    payload[key].clear()
                torch.save(payload, fname)
                return
            except Exception:
                if attempt:
                    raise
                sleep(np.random.rand() * 5 + 2)
