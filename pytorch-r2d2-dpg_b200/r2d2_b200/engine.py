"""Host-side orchestration of the native learner path (torch = device memory, streams, NCCL).

`LearnerEngine` owns the flat parameter / gradient / Adam buffers (torch CUDA tensors), exposes them
as `state_dict()`-compatible views with the reference's keys (models.py:17-19), and drives the three
native phases of one learner iteration (learner.py:84-139).  With torch.distributed initialised it
all-reduces the two flat gradient buffers between the phases: nothing else crosses GPUs.

`DeviceReplay` is the per-GPU replay shard (replaces LearnerReplayMemory storage + sampling,
replay_memory.py:67-136) with the sum tree in HBM.
"""
from __future__ import annotations

from collections import OrderedDict
from ctypes import byref, c_int, c_longlong, c_void_p
from dataclasses import dataclass

import numpy as np
import torch

from . import native as nv

PARAM_KEYS = ("l1.weight", "l1.bias", "l2.weight_ih", "l2.weight_hh", "l2.bias_ih", "l2.bias_hh",
              "l3.weight", "l3.bias")


@dataclass
class PathConfig:
    """Hyper-parameters of the path; defaults are the reference's literals (learner.py:29-34,43-49)."""
    obs: int
    act: int
    hidden: int = 128
    batch: int = 32
    burn_in: int = 20
    learning: int = 40
    n_step: int = 5
    gamma: float = 0.997
    actor_lr: float = 1e-4
    critic_lr: float = 1e-3
    eta: float = 0.9
    target_interval: int = 500

    @property
    def rows(self) -> int:
        return self.burn_in + self.learning + self.n_step


def param_shapes(cfg: PathConfig, critic: bool):
    H, A = cfg.hidden, cfg.act
    I = cfg.obs + (cfg.act if critic else 0)
    return OrderedDict([("l1.weight", (H, I)), ("l1.bias", (H,)), ("l2.weight_ih", (4 * H, H)),
                        ("l2.weight_hh", (4 * H, H)), ("l2.bias_ih", (4 * H,)), ("l2.bias_hh", (4 * H,)),
                        ("l3.weight", (A, H)), ("l3.bias", (A,))])


def flat_views(flat: torch.Tensor, cfg: PathConfig, critic: bool):
    """state_dict-ordered views into a flat parameter block (no copies)."""
    out, off = OrderedDict(), 0
    for k, shp in param_shapes(cfg, critic).items():
        n = int(np.prod(shp))
        out[k] = flat[off:off + n].view(shp)
        off += n
    assert off == flat.numel()
    return out


def init_reference_params(cfg: PathConfig, critic: bool, generator: torch.Generator | None = None):
    """Fresh parameters with the reference's distributions (models.py:8-11,21-25): fan-in uniform keyed
    on out_features for l1 / LSTM weights, LSTMCell-default U(+-1/sqrt(H)) biases, Linear-default l1 bias,
    l3 U(+-3e-3) / 3e-4.  (Seed-for-seed identity with torch's module constructors is not needed here;
    parity tests load the reference's own tensors.)"""
    H = cfg.hidden
    shapes = param_shapes(cfg, critic)
    u = lambda shp, b: (torch.rand(shp, generator=generator) * 2 - 1) * b  # noqa: E731
    I = shapes["l1.weight"][1]
    p = OrderedDict()
    p["l1.weight"] = u(shapes["l1.weight"], 1.0 / np.sqrt(H))
    p["l1.bias"] = u(shapes["l1.bias"], 1.0 / np.sqrt(I))
    p["l2.weight_ih"] = u(shapes["l2.weight_ih"], 1.0 / np.sqrt(4 * H))
    p["l2.weight_hh"] = u(shapes["l2.weight_hh"], 1.0 / np.sqrt(4 * H))
    p["l2.bias_ih"] = u(shapes["l2.bias_ih"], 1.0 / np.sqrt(H))
    p["l2.bias_hh"] = u(shapes["l2.bias_hh"], 1.0 / np.sqrt(H))
    p["l3.weight"] = u(shapes["l3.weight"], 3e-3)
    p["l3.bias"] = torch.full(shapes["l3.bias"], 3e-4)
    return p


class LearnerEngine:
    def __init__(self, cfg: PathConfig, device=None, seed: int = 1):
        if not torch.cuda.is_available():
            raise nv.NativeError("LearnerEngine needs a CUDA device (B200); there is no CPU fallback")
        self.lib = nv.lib()
        self.cfg = cfg
        self._pending_finish = False          # a deferred phase 3 (data-parallel "defer" mode), see step() / flush()
        self._h = None
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        na = sum(int(np.prod(s)) for s in param_shapes(cfg, False).values())
        nc = sum(int(np.prod(s)) for s in param_shapes(cfg, True).values())
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)  # noqa: E731
        self.flat = {"actor": z(na), "critic": z(nc), "target_actor": z(na), "target_critic": z(nc)}
        self.grads = {"actor": z(na), "critic": z(nc)}
        self.exp_avg = {"actor": z(na), "critic": z(nc)}
        self.exp_avg_sq = {"actor": z(na), "critic": z(nc)}
        g = torch.Generator().manual_seed(seed)
        self.load_state_dicts(init_reference_params(cfg, False, g), init_reference_params(cfg, True, g))
        c = nv.LearnerConfig(cfg.obs, cfg.act, cfg.hidden, cfg.batch, cfg.burn_in, cfg.learning, cfg.n_step,
                             cfg.gamma, cfg.actor_lr, cfg.critic_lr, cfg.eta, cfg.target_interval,
                             self.flat["actor"].data_ptr(), self.flat["critic"].data_ptr(),
                             self.flat["target_actor"].data_ptr(), self.flat["target_critic"].data_ptr(),
                             self.grads["actor"].data_ptr(), self.grads["critic"].data_ptr(),
                             self.exp_avg["actor"].data_ptr(), self.exp_avg_sq["actor"].data_ptr(),
                             self.exp_avg["critic"].data_ptr(), self.exp_avg_sq["critic"].data_ptr())
        self._h = c_void_p()
        nv.check(self.lib.r2d2_learner_create(byref(self._h), byref(c)))
        T, B, O, A, H, L = cfg.rows, cfg.batch, cfg.obs, cfg.act, cfg.hidden, cfg.learning
        dev = self.device
        # the batch has two slots (include/r2d2_b200.h): the attributes obs .. uniforms are views of the slot the NEXT
        # sample goes to; a plain caller never leaves slot 0
        self._slots = []
        for slot in (0, 1):
            b = nv.LearnerBuffers()
            nv.check(self.lib.r2d2_learner_buffers_get_slot(self._h, slot, byref(b)))
            self._slots.append({"obs": nv.view_f32(b.obs, (T, B, O), dev), "act": nv.view_f32(b.act, (T, B, A), dev),
                                "rew": nv.view_f32(b.rew, (T, B), dev), "term": nv.view_f32(b.term, (T, B), dev),
                                "states": nv.view_f32(b.states, (4, 2, B, H), dev),
                                "leaf_idx": nv.view_i64(b.leaf_idx, (B,), dev),
                                "uniforms": nv.view_f32(b.uniforms, (B,), dev)})
        self._lib_slot = 0
        self._targets_ahead = False      # the fill slot's target chains already ran: its batch must not change any more
        self._bind_slot(0)
        self.q_value = nv.view_f32(b.q_value, (L * B, A), dev)
        self.target_q_value = nv.view_f32(b.target_q_value, (L * B, A), dev)
        self.td_sq = nv.view_f32(b.td_sq, (L * B,), dev)
        self.priority = nv.view_f32(b.priority, (B,), dev)
        self.losses = nv.view_f32(b.losses, (2,), dev)
        self.world = 1
        self._dist = None
        self._sync = None
        import os
        # data-parallel gradient exchange: "peer" (default: the library's own kernels over NVLink peer memory, in the
        # learner's stream), "defer" (NCCL all-reduces on a side stream, the actor's waited for one critic phase later;
        # also the fallback when no peer-mapped buffer can be set up), A/B timing only: "overlap" (NCCL, both waited for
        # where needed), "serial" (NCCL on the compute stream), "none" (no exchange at all: replicas diverge)
        self._dp_mode = os.environ.get("R2D2_DP_MODE", "peer")
        self._peer_buf = None
        self._peer_hdl = None
        self._pending_finish = False
        self._sync_actor = None

    def _guard_fill(self):
        if self._targets_ahead:
            raise nv.NativeError("the engine's batch was filled by a step(prefetch=...) hook and its target chains "
                                 "have already run: call step(), or discard_prefetched(), before writing another batch")

    def discard_prefetched(self):
        """Drop a batch that a step(prefetch=...) hook drew ahead (its target chains are forgotten too): the next
        sample_into / set_batch starts a fresh sequence."""
        nv.check(self.lib.r2d2_learner_discard_prefetch(self._h, nv.current_stream()))
        self._targets_ahead = False

    def _bind_slot(self, slot: int):
        for k, v in self._slots[slot].items():
            setattr(self, k, v)
        self._fill_slot = slot

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            if getattr(self, "_peer_buf", None) is not None:
                # peers read this rank's gradient block until their last slice sum ran: completing the pending phase 3
                # (its wait kernel needs every peer's "slice delivered" flag) proves nobody touches the buffer any more
                try:
                    self.flush()
                    torch.cuda.synchronize(self.device)
                except Exception:
                    pass
            self._pending_finish = False
            self.lib.r2d2_learner_destroy(self._h)
            self._h = None
            self._peer_buf = self._peer_hdl = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------
    def views(self, net: str, what: str = "params"):
        self.flush()
        src = {"params": self.flat, "grads": self.grads, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}[what]
        return flat_views(src[net], self.cfg, "critic" in net)

    def state_dict(self, net: str):
        return OrderedDict((k, v.detach().clone()) for k, v in self.views(net).items())

    def load_state_dicts(self, actor, critic, target_actor=None, target_critic=None):
        def put(net, sd):
            for k, v in self.views(net).items():
                v.copy_(torch.as_tensor(np.asarray(sd[k]) if not isinstance(sd[k], torch.Tensor) else sd[k],
                                        dtype=torch.float32).to(self.device))
        put("actor", actor)
        put("critic", critic)
        put("target_actor", target_actor if target_actor is not None else actor)
        put("target_critic", target_critic if target_critic is not None else critic)

    def enable_data_parallel(self, require: bool | None = None):
        """Gradients are averaged over ranks at the two optimiser steps: by the library's own kernels over NVLink peer
        memory (`_attach_peers`, csrc/peer.cu), or - fallback / R2D2_DP_MODE=defer - by NCCL all-reduces of the flat
        buffers on a side stream.  `require` (default: WORLD_SIZE > 1 in the environment) turns a missing process group
        into an error instead of N silently independent learners."""
        import os
        import torch.distributed as dist
        from .dist_env import GradSync
        if require is None:
            require = int(os.environ.get("WORLD_SIZE", "1")) > 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._dist = dist
            self.world = dist.get_world_size()
            self._sync = GradSync(dist, self.world)
            self._sync_actor = GradSync(dist, self.world)
            if self._dp_mode == "peer":
                self._attach_peers(dist)
            if self._dp_mode in ("peer", "defer"):   # the actor's weights are final only after the deferred phase 3
                nv.check(self.lib.r2d2_learner_set_overlap_actor_inputs(self._h, 0))
            for net in ("actor", "critic", "target_actor", "target_critic"):
                dist.broadcast(self.flat[net], src=0)
            for d in (self.exp_avg, self.exp_avg_sq):
                for net in ("actor", "critic"):
                    dist.broadcast(d[net], src=0)
        elif require:
            raise nv.NativeError("WORLD_SIZE > 1 but torch.distributed is not initialised: call "
                                 "r2d2_b200.dist_env.DistEnv.from_environ().init_process_group() first "
                                 "(the drop-in learner.Learner does)")

    def _attach_peers(self, dist):
        """Move the gradient blocks into a buffer every rank of the node maps (torch symmetric memory: CUDA fabric /
        IPC handles exchanged through the process group's store) and hand the peer addresses to the library.  All ranks
        agree on the outcome; if any rank cannot map its peers, every rank falls back to the NCCL "defer" mode."""
        ok, why = 1, ""
        try:
            import torch.distributed._symmetric_memory as symm
            lay = nv.PeerLayout()
            nv.check(self.lib.r2d2_learner_peer_layout(self._h, self.world, byref(lay)))
            buf = symm.empty(int(lay.bytes) // 4, dtype=torch.float32, device=self.device)
            buf.zero_()
            hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or int(hdl.rank) != dist.get_rank() or ptrs[dist.get_rank()] != buf.data_ptr():
                raise RuntimeError("symmetric-memory handle does not match the process group")
        except Exception as e:  # noqa: BLE001 - any failure means "no peer mapping on this box"
            ok, why = 0, repr(e)[:200]
        agreed = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()) == 0:
            if dist.get_rank() == 0:
                print(f"[r2d2_b200] no peer-mapped gradient buffer ({why or 'another rank failed'}): "
                      "gradient exchange falls back to NCCL all-reduces (R2D2_DP_MODE=defer)", flush=True)
            self._dp_mode = "defer"
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()                                   # every rank's flag words are zero before anyone signals
        arr = (c_void_p * self.world)(*ptrs)
        nv.check(self.lib.r2d2_learner_attach_peers(self._h, dist.get_rank(), self.world, arr))
        self._peer_buf, self._peer_hdl = buf, hdl
        na, nc = self.grads["actor"].numel(), self.grads["critic"].numel()
        oc, oa = int(lay.off_critic_grads) // 4, int(lay.off_actor_grads) // 4
        self.grads = {"actor": buf[oa:oa + na], "critic": buf[oc:oc + nc]}

    def peer_status(self) -> int:
        """0, or 1 when a bounded wait for a peer's flag expired inside the gradient-exchange kernels."""
        if self._peer_buf is None:
            return 0
        st = c_int(0)
        nv.check(self.lib.r2d2_learner_peer_status(self._h, byref(st), nv.current_stream()))
        return int(st.value)

    def replicas_identical(self) -> bool:
        """Data-parallel invariant: parameters and Adam moments are bit-identical on every rank."""
        if self._dist is None:
            return True
        self.flush()
        if self.peer_status() != 0:
            return False
        ts = [self.flat[n] for n in ("actor", "critic", "target_actor", "target_critic")]
        ts += [d[n] for d in (self.exp_avg, self.exp_avg_sq) for n in ("actor", "critic")]
        return self._sync.replicas_identical(ts)

    # ---- batch ------------------------------------------------------------------------------
    def set_batch(self, batch: dict):
        """Copy an already sampled time-major batch (replay_memory.py:123-136 layout) into the engine."""
        cv = lambda x: torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x,  # noqa: E731
                                       dtype=torch.float32).to(self.device, non_blocking=True)
        self._guard_fill()
        self.obs.copy_(cv(batch["obs"]))
        self.act.copy_(cv(batch["act"]))
        self.rew.copy_(cv(batch["rew"]).reshape(self.rew.shape))
        self.term.copy_(cv(batch["term"]).reshape(self.term.shape))
        for i, k in enumerate(("a_state", "ta_state", "c_state", "tc_state")):
            self.states[i].copy_(cv(batch[k]))

    # ---- one learner iteration (learner.py:86-132) on the batch currently in the engine ------------
    def step(self, prefetch=None):
        """One learner iteration on the batch in the engine (learner.py:86-132).

        `prefetch(engine, used)`, if given, is called exactly once per step and must (1) write back the priorities of
        the batch this step trained on - `used.leaf_idx`, `used.priority` - and (2) fill the engine's batch buffers
        (`obs .. states`, e.g. `DeviceReplay.sample_into`) with the NEXT batch.  It is called as early as the data flow
        allows - right after the critic phase, when the priorities exist - and the next batch's target chains run
        straight away, in the middle of this iteration: they read only the target nets, so they are the independent
        work that lets the data-parallel ranks drift by a millisecond or two without waiting for each other (and on
        one GPU the input projections hide under their scans as before).  On iterations whose finish phase copies
        into the target nets it is called at the end of the step instead.

        Data parallel, mode "peer" (default): the gradient blocks are summed by the library's own kernels over NVLink
        peer memory inside the phases (csrc/peer.cuh); the actor's optimiser step of iteration i runs after the critic
        phase of iteration i+1.  Mode "defer" (fallback): NCCL all-reduces on a side stream, the actor's waited for
        after the next critic phase; the hook then runs at the end of the step."""
        s = nv.current_stream()
        scale = 1.0 / self.world
        mode = self._dp_mode if self._dist is not None else "single"
        if self._fill_slot != self._lib_slot:
            nv.check(self.lib.r2d2_learner_select_batch(self._h, self._fill_slot))
            self._lib_slot = self._fill_slot
        self._targets_ahead = False
        if self._pending_finish and self._finish_updates_targets():
            self.flush()                                                  # the target chains below read the target nets
        nv.check(self.lib.r2d2_learner_critic_phase(self._h, s))
        if mode in ("peer", "single", "none"):
            if mode == "peer":
                self.flush()                                              # phase 3 of the previous iteration
            ahead = prefetch is not None and not self._finish_updates_targets()
            if ahead:
                self._run_prefetch(prefetch)
                nv.check(self.lib.r2d2_learner_target_phase(self._h, self._fill_slot, s))
                self._targets_ahead = True
            nv.check(self.lib.r2d2_learner_actor_forward(self._h, s))
            nv.check(self.lib.r2d2_learner_actor_phase(self._h, scale, s))
            if mode == "peer":   # signal / slice-sum / wait kernels are issued by the phases themselves
                self._pending_finish = True
            else:
                nv.check(self.lib.r2d2_learner_finish_phase(self._h, scale, s))
            if prefetch is not None and not ahead:
                self._run_prefetch(prefetch)
            return
        if mode in ("defer", "overlap"):
            self._sync.start(self.grads["critic"])                       # side stream
            self.flush()                                                  # phase 3 of the previous iteration (actor Adam)
            nv.check(self.lib.r2d2_learner_actor_forward(self._h, s))    # reads no critic weights: overlaps the all-reduce
            self._sync.wait(self.device)
        elif mode == "serial":                                            # A/B: both all-reduces on the compute stream
            self._dist.all_reduce(self.grads["critic"])
        nv.check(self.lib.r2d2_learner_actor_phase(self._h, scale, s))
        if mode == "defer":
            self._sync_actor.start(self.grads["actor"])
            self._pending_finish = True
        else:
            if mode == "overlap":
                self._sync_actor.start(self.grads["actor"])
                self._sync_actor.wait(self.device)
            elif mode == "serial":
                self._dist.all_reduce(self.grads["actor"])
            nv.check(self.lib.r2d2_learner_finish_phase(self._h, scale, s))
        if prefetch is not None:
            self._run_prefetch(prefetch)

    def _run_prefetch(self, prefetch):
        from types import SimpleNamespace
        used = SimpleNamespace(leaf_idx=self.leaf_idx, priority=self.priority, losses=self.losses)
        self._bind_slot(1 - self._fill_slot)     # the phases still in flight keep reading the other slot
        prefetch(self, used)

    def _finish_updates_targets(self) -> bool:
        k = self.cfg.target_interval
        return k > 0 and (int(self.lib.r2d2_learner_step_count(self._h)) + 1) % k == 0

    def flush(self):
        """Complete a deferred phase 3 (actor all-reduce wait + Adam + step counter + target update)."""
        if self._pending_finish:
            if self._dp_mode != "peer":
                self._sync_actor.wait(self.device)
            nv.check(self.lib.r2d2_learner_finish_phase(self._h, 1.0 / self.world, nv.current_stream()))
            self._pending_finish = False

    @property
    def step_count(self) -> int:
        self.flush()
        return int(self.lib.r2d2_learner_step_count(self._h))

    # ---- full training state (the reference checkpoints weights only and cannot resume) ---------------
    def training_state(self) -> dict:
        """Everything a restart needs: the four nets (reference keys), both Adam moment sets, the step counter."""
        torch.cuda.synchronize(self.device)
        out = {net: self.state_dict(net) for net in ("actor", "target_actor", "critic", "target_critic")}
        for net in ("actor", "critic"):
            out[net + "_optimizer"] = {"exp_avg": OrderedDict((k, v.detach().clone()) for k, v in self.views(net, "exp_avg").items()),
                                       "exp_avg_sq": OrderedDict((k, v.detach().clone()) for k, v in self.views(net, "exp_avg_sq").items())}
        out["step"] = self.step_count
        return out

    def load_training_state(self, st: dict):
        self.load_state_dicts(st["actor"], st["critic"], st.get("target_actor"), st.get("target_critic"))
        for net in ("actor", "critic"):
            opt = st.get(net + "_optimizer")
            if opt is not None:
                for what in ("exp_avg", "exp_avg_sq"):
                    for k, v in self.views(net, what).items():
                        v.copy_(torch.as_tensor(opt[what][k], dtype=torch.float32).to(self.device))
        nv.check(self.lib.r2d2_learner_set_step_count(self._h, int(st.get("step", 0))))

    @property
    def launches_per_iteration(self) -> int:
        return int(self.lib.r2d2_learner_launches_per_iteration(self._h))


class DeviceReplay:
    """One replay shard in HBM with a 32-ary sum tree (one leaf per stored row)."""

    def __init__(self, cfg: PathConfig, capacity_rows: int, max_sequences: int = 0, device=None):
        if not torch.cuda.is_available():
            raise nv.NativeError("DeviceReplay needs a CUDA device; there is no CPU fallback")
        self.lib = nv.lib()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        rc = nv.ReplayConfig(cfg.obs, cfg.act, cfg.hidden, cfg.burn_in, cfg.learning, cfg.n_step,
                             int(capacity_rows), int(max_sequences))
        self._h = c_void_p()
        nv.check(self.lib.r2d2_replay_create(byref(self._h), byref(rc)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.r2d2_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_episode(self, obs, act, rew, term, states, priority):
        """obs [n_rows,O], act [n_rows,A], rew/term [n_rows], states [n_real,4,2,H], priority [n_starts] (host)."""
        obs, po = nv.host_f32(obs)
        act, pa = nv.host_f32(act)
        rew, pr = nv.host_f32(np.asarray(rew).reshape(-1))
        term, pt = nv.host_f32(np.asarray(term).reshape(-1))
        states, ps = nv.host_f32(states)
        priority, pp = nv.host_f32(np.asarray(priority).reshape(-1))
        nv.check(self.lib.r2d2_replay_add_episode(self._h, po, pa, pr, pt, ps, obs.shape[0], states.shape[0], pp,
                                                  priority.shape[0], nv.current_stream()))

    def add_episodes(self, episodes):
        """One actor file in one native call (LearnerReplayMemory.load, replay_memory.py:138-157).  `episodes`: list of
        (obs [n,O], act [n,A], rew [n], term [n], states [n_real,4,2,H], priority [n_starts]) host arrays.  Returns
        (row_start per episode, episodes evicted by the call, sequence counter)."""
        if not episodes:
            return [], 0, None
        n_rows = np.asarray([e[0].shape[0] for e in episodes], np.int32)
        n_starts = np.asarray([len(e[5]) for e in episodes], np.int32)
        R, H = int(n_rows.sum()), self.cfg.hidden
        obs = np.concatenate([np.asarray(e[0], np.float32) for e in episodes])
        act = np.concatenate([np.asarray(e[1], np.float32) for e in episodes])
        rew = np.concatenate([np.asarray(e[2], np.float32).reshape(-1) for e in episodes])
        term = np.concatenate([np.asarray(e[3], np.float32).reshape(-1) for e in episodes])
        states = np.zeros((R, 4, 2, H), np.float32)
        leaf = np.zeros(R, np.float32)
        off = 0
        for e, n in zip(episodes, n_rows):
            st = np.asarray(e[4], np.float32)
            if st.shape[1:] != (4, 2, H) or st.shape[0] > n:
                raise ValueError("recurrent states %r do not fit an episode of %d rows at hidden %d" % (st.shape, n, H))
            states[off:off + st.shape[0]] = st
            leaf[off:off + len(e[5])] = np.asarray(e[5], np.float32).reshape(-1)
            off += int(n)
        if obs.shape != (R, self.cfg.obs) or act.shape != (R, self.cfg.act):
            raise ValueError("episode rows %r / %r do not match the shard (obs %d, act %d)" % (obs.shape, act.shape,
                                                                                           self.cfg.obs, self.cfg.act))
        starts = np.zeros(len(episodes), np.int64)
        n_evicted, counter = c_longlong(0), c_longlong(0)
        P = lambda a: a.ctypes.data_as(c_void_p)  # noqa: E731
        nv.check(self.lib.r2d2_replay_add_episodes(self._h, len(episodes), P(n_rows), P(n_starts), P(obs), P(act), P(rew),
                                                   P(term), P(states), P(leaf), P(starts), byref(n_evicted), byref(counter),
                                                   nv.current_stream()))
        return starts.tolist(), int(n_evicted.value), int(counter.value)

    def sample_indices(self, u: torch.Tensor) -> torch.Tensor:
        leaf = torch.empty(u.numel(), dtype=torch.int64, device=self.device)
        nv.check(self.lib.r2d2_replay_sample(self._h, nv.dptr(u), u.numel(), nv.dptr(leaf, torch.int64), None, None,
                                             None, None, None, nv.current_stream()))
        return leaf

    def sample_into(self, eng: LearnerEngine, generator: torch.Generator | None = None, u: torch.Tensor | None = None):
        """Draw eng.cfg.batch starts and gather the time-major batch straight into the engine's buffers."""
        ec, rc = eng.cfg, self.cfg
        eng._guard_fill()
        if (ec.obs, ec.act, ec.hidden, ec.rows) != (rc.obs, rc.act, rc.hidden, rc.rows):
            raise nv.NativeError("replay shard (obs %d act %d hidden %d rows %d) does not match the engine (obs %d act %d "
                                 "hidden %d rows %d)" % (rc.obs, rc.act, rc.hidden, rc.rows, ec.obs, ec.act, ec.hidden, ec.rows))
        if u is None:
            eng.uniforms.copy_(torch.rand(eng.cfg.batch, device=self.device, generator=generator))
        else:
            eng.uniforms.copy_(u)
        nv.check(self.lib.r2d2_replay_sample(self._h, nv.dptr(eng.uniforms), eng.cfg.batch,
                                             nv.dptr(eng.leaf_idx, torch.int64), nv.dptr(eng.obs), nv.dptr(eng.act),
                                             nv.dptr(eng.rew), nv.dptr(eng.term), nv.dptr(eng.states),
                                             nv.current_stream()))

    def update_priorities(self, leaf_idx: torch.Tensor, prio: torch.Tensor):
        nv.check(self.lib.r2d2_replay_update_priorities(self._h, nv.dptr(leaf_idx, torch.int64), nv.dptr(prio),
                                                        leaf_idx.numel(), nv.current_stream()))

    def stats(self) -> dict:
        st = nv.ReplayStats()
        nv.check(self.lib.r2d2_replay_stats(self._h, byref(st), nv.current_stream()))
        return {k: getattr(st, k) for k, _ in nv.ReplayStats._fields_}

    def decode(self, leaf_idx):
        leaf = np.ascontiguousarray(np.asarray(leaf_idx.cpu() if isinstance(leaf_idx, torch.Tensor) else leaf_idx),
                                    dtype=np.int64)
        ep = np.empty_like(leaf)
        sq = np.empty_like(leaf)
        nv.check(self.lib.r2d2_replay_decode(self._h, leaf.ctypes.data_as(c_void_p), leaf.size,
                                             ep.ctypes.data_as(c_void_p), sq.ctypes.data_as(c_void_p)))
        return ep, sq

    def tree_level(self, level: int) -> torch.Tensor:
        p, n = c_void_p(), c_longlong()
        nv.check(self.lib.r2d2_replay_tree_level(self._h, level, byref(p), byref(n)))
        return nv.view_f32(p.value, (n.value,), self.device)
