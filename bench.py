#!/usr/bin/env python
"""bench.py - learner sequence-steps/sec (batch x seq_len per learner iteration) of the B200-native
learner hot path, BASELINE.json configs[1]: synthetic obs=17 act=6 hidden=256 seq_len=80 burn_in=40
batch=256 per GPU (weak scaling: every rank owns a replay shard and a batch of 256; gradients are
all-reduced over NCCL at the two optimiser steps).

  python bench.py --gpus N --steps K --warmup W          # torchrun launches one process per GPU for N > 1
  python bench.py --impl reference ...                   # the reference's CPU implementation (oracle port)

A step = one pass of learner.py:84-139: prioritized sample from the HBM replay shard -> gather ->
target/online chains -> TD/priority kernel -> critic BPTT + Adam -> actor chain -> DPG backward + Adam
-> priority write-back into the sum tree.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch-r2d2-dpg_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0] shapes (reference as-is, walker sizes) / configs[1] (headline) / configs[2]
    "cfg1": dict(obs=24, act=6, hidden=128, batch=32, burn_in=20, learning=40, n_step=5),
    "cfg2": dict(obs=17, act=6, hidden=256, batch=256, burn_in=40, learning=80, n_step=5),
    "cfg3": dict(obs=376, act=17, hidden=512, batch=512, burn_in=40, learning=80, n_step=5),
}
METRIC = "learner sequence-steps/sec (batch x seq_len)"


def lstm_flops_per_iteration(c):
    """FLOP_lstm = 16*B*H^2*(5*Bn + 13*L + 2*n) (necessary cell-steps, fwd + bwd)."""
    return 16.0 * c["batch"] * c["hidden"] ** 2 * (5 * c["burn_in"] + 13 * c["learning"] + 2 * c["n_step"])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self, begin):
        """wall-clock window of the timed region: only samples taken inside it are reported"""
        if begin:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.05]
        rows = inside if inside else [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_replay(engine, cfg, n_episodes, episode_len, seed, device):
    rng = np.random.default_rng(seed)
    n_rows = episode_len + cfg.n_step
    rp = engine.DeviceReplay(cfg, capacity_rows=n_episodes * n_rows, device=device)
    for _ in range(n_episodes):
        term = np.zeros(n_rows, np.float32)
        term[episode_len:] = 1
        obs = rng.standard_normal((n_rows, cfg.obs), dtype=np.float32)
        act = rng.uniform(-1, 1, (n_rows, cfg.act)).astype(np.float32)
        rew = rng.standard_normal(n_rows, dtype=np.float32)
        obs[episode_len:] = 0
        act[episode_len:] = 0
        rew[episode_len:] = 0
        states = 0.1 * rng.standard_normal((episode_len, 4, 2, cfg.hidden), dtype=np.float32)
        prio = rng.uniform(0.01, 1.0, episode_len - (cfg.burn_in + cfg.learning)).astype(np.float32)
        rp.add_episode(obs, act, rew, term, states, prio)
    return rp


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def cpu_model():
    """CPU model string of the host the CPU baseline ran on (SURVEY 8d asks for it next to os.cpu_count())."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def pick_cpu_threads(c):
    """torch CPU ops of this size (batch x hidden LSTMCell steps) stop scaling - and can collapse - long
    before a 100+ core box is full, so probe a few thread counts on a short LSTMCell fwd+bwd loop and keep
    the fastest: the baseline gets the best setting this host offers, and `cores` reports it."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (4, 8, 16, 32, 64, ncpu) if t <= ncpu})
    cell = torch.nn.LSTMCell(c["hidden"], c["hidden"])
    x = torch.randn(c["batch"], c["hidden"])
    best, best_t = cands[0], float("inf")
    for t in cands:
        torch.set_num_threads(t)
        for rep in range(2):
            t0 = time.perf_counter()
            h = cx = torch.zeros(c["batch"], c["hidden"])
            for _ in range(12):
                h, cx = cell(x, (h, cx))
            h.sum().backward()
            dt = time.perf_counter() - t0
        log(f"cpu thread probe: {t} threads -> {dt * 1e3:.1f} ms")
        if dt < best_t:
            best, best_t = t, dt
        elif dt > 2.0 * best_t:
            break  # past the knee: more threads only add spin-wait contention (128 threads: 1800x slower here)
    return best


def time_cpu_port(c, steps, warmup, threads=None, budget_s=150.0):
    """The reference's CPU implementation of the path (oracle/ref_port.py: same torch CPU operators,
    python loops, autograd, Adam, two-level WeightedRandomSampler draw) on this box's host cores.
    Stops early (after >= 1 timed iteration) when `budget_s` of wall clock is spent."""
    from oracle import ref_port
    threads = threads or pick_cpu_threads(c)
    torch.set_num_threads(threads)
    pc = ref_port.PathConfig(**c)
    lr = ref_port.PortLearner(pc, seed=1)
    rp = ref_port.synthetic_replay(pc, n_episodes=max(8, (2 * c["batch"]) // 100 + 8), episode_len=250, seed=0)
    times, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ep, sq, batch = rp.sample()
        out = lr.iteration(batch, keep_tensors=False)
        rp.write_back(ep, sq, out["priority"])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        log(f"cpu port iteration {i}: {time.perf_counter() - t0:.2f} s")
        if times and time.perf_counter() - t_start > budget_s:
            break
    sec = float(np.mean(times))
    return c["batch"] * c["learning"] / sec, sec, threads, len(times)


def run_reference(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, threads, done = time_cpu_port(c, args.steps, args.warmup, budget_s=240.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "seq-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "steps_timed": done, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: " + " ".join(f"{k}={v}" for k, v in c.items()),
                       "note": "reference CPU learner (oracle/ref_port.py port; /root/reference is python and not "
                               "present on this box), full batch per step; a step is one full learner iteration "
                               "(~1-2 s on the host): at most 240 s of them are timed (steps_timed), the rate does "
                               "not depend on the count"},
            "cpu_baseline": {"value": value, "unit": "seq-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{done} full learner iterations at batch {c['batch']} after {args.warmup} warm-up, "
                                       f"{threads} torch threads (best of a probe over {os.cpu_count()} host cores)",
                             "cpu_model": cpu_model(), "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": "seq-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: point fd 1 at stderr for everything libraries print (NCCL / c10d
    banners go to fd 1 of every rank) and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--episodes", type=int, default=384, help="episodes in the per-GPU replay shard")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    args = ap.parse_args()
    c = CONFIGS[args.config]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, c)

    from r2d2_b200 import engine, native as nv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: keep NCCL's banner / debug output out of it
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_bench_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    log(f"world={world} rank={rank}: building engine + replay shard")
    cfg = engine.PathConfig(**c)
    eng = engine.LearnerEngine(cfg, device=dev, seed=1)
    eng.enable_data_parallel()
    ep_len = 250
    rp = build_replay(engine, cfg, args.episodes, ep_len, seed=100 + rank, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    B, L, T, O, A, H = cfg.batch, cfg.learning, cfg.rows, cfg.obs, cfg.act, cfg.hidden

    def step_resident():
        rp.sample_into(eng, generator=gen)
        eng.step()
        rp.update_priorities(eng.leaf_idx, eng.priority)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    log("warm-up (HBM-resident arm)")
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # started before the warm-up so that samples exist when the timed region begins
    for _ in range(args.warmup):
        step_resident()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark(True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    clocks.mark(False)
    ms = ev0.elapsed_time(ev1) / args.steps
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * B * L / (ms * 1e-3)
    launches_per_step = eng.launches_per_iteration + 6 + 1   # + sample/gather kernels + tree update

    log(f"resident arm: {ms:.3f} ms/step; e2e arm")
    # ---- e2e: the same iteration fed from HOST buffers (the reference's boundary: replay_memory.py:123-133
    # copies the sampled batch to the device each iteration, learner.py:135 reads the TD result back)
    n_pool = 4
    pool = []
    for i in range(n_pool):
        rp.sample_into(eng, generator=gen)
        torch.cuda.synchronize()
        pool.append({k: getattr(eng, k).cpu().pin_memory() for k in ("obs", "act", "rew", "term", "states")})
    host_prio = torch.empty(B, dtype=torch.float32).pin_memory()
    host_loss = torch.empty(2, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * 4 for v in pool[0].values())
    d2h = (B + 2) * 4

    # Double-buffered feed: the pinned batch of step i+1 crosses PCIe on a copy stream while step i computes; the
    # compute stream only does a device-to-device move of the staged batch (5 MB at HBM speed) before each step.
    keys = list(pool[0].keys())
    copy_stream = torch.cuda.Stream()
    stage = [{k: torch.empty_like(getattr(eng, k)) for k in keys} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for e in consumed:
        e.record(torch.cuda.current_stream())

    def prefetch(i):
        s_ = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s_])          # the staged batch of step i-2 has been moved out
            for k, v in pool[i % n_pool].items():
                stage[s_][k].copy_(v, non_blocking=True)
            ready[s_].record(copy_stream)

    def step_host(i):
        s_ = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[s_])
        for k in keys:
            getattr(eng, k).copy_(stage[s_][k], non_blocking=True)
        consumed[s_].record(cur)
        prefetch(i + 1)                                   # H2D of the next step's inputs, overlapped with this step
        eng.step()
        host_prio.copy_(eng.priority, non_blocking=True)
        host_loss.copy_(eng.losses, non_blocking=True)
        cur.synchronize()                                 # the host consumes the priorities every iteration

    prefetch(0)
    for i in range(args.warmup):
        step_host(i)
    barrier()
    ev0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        step_host(i)
    ev1.record()
    barrier()
    copy_stream.synchronize()
    ms_e2e = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    if dist is not None:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * L / (float(ms_e2e.item()) * 1e-3)

    log("roofline: timing the scan kernel alone")
    # ---- roofline of the dominant kernel: the persistent LSTM scan (serial half of every cell step).
    # algorithmic FLOPs per launch = 2 * B * H * 4H per cell step x S steps (F_cell = 16 B H^2 covers
    # both halves; the hoisted x*W_ih half runs in gemm_f32).  Timed alone with CUDA events on the launch stream.
    peaks = measured_peaks()
    S = cfg.burn_in + cfg.n_step + cfg.learning
    gin = torch.randn(S, B, 4 * H, device=dev) * 0.5
    whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / np.sqrt(4 * H)
    gates = torch.empty_like(gin)
    hs = torch.empty(S + 1, B, H, device=dev)
    cs = torch.empty(S + 1, B, H, device=dev)
    lib, st = nv.lib(), nv.current_stream()
    scratch = torch.empty(B * 4 * H, device=dev)

    def scan():
        nv.check(lib.r2d2_lstm_scan_forward(nv.dptr(gin), nv.dptr(whh), None, None, nv.dptr(gates), nv.dptr(hs),
                                            nv.dptr(cs), None, S, B, H, 1, nv.dptr(scratch), st))
    for _ in range(3):
        scan()
    torch.cuda.synchronize()
    reps = 10
    ev0.record()
    for _ in range(reps):
        scan()
    ev1.record()
    torch.cuda.synchronize()
    scan_ms = ev0.elapsed_time(ev1) / reps
    scan_flops = 2.0 * B * H * 4 * H * S
    achieved = scan_flops / (scan_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scan_fwd_traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "lstm_scan_fwd_pp_kernel (persistent cluster LSTM scan, tcgen05 ping-pong over two row sub-tiles, W_hh in TMEM, %d cell steps)" % S, "bound": "tensor",
                "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"],
                "traffic": traffic, "peak_source": peaks["source"] + " bf16 dense burst (kernel timed alone)",
                "us_per_step": scan_ms * 1e3 / S,
                "whole_iteration": {"lstm_flops": lstm_flops_per_iteration(c),
                                    "achieved_tflops": lstm_flops_per_iteration(c) / (ms * 1e-3) / 1e12,
                                    "frac_of_sustained_peak": lstm_flops_per_iteration(c) / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"]}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        log("cpu_baseline leg (oracle port on host cores)")
        v, sec, threads, done = time_cpu_port(c, args.cpu_steps, 1, budget_s=60.0)
        cpu_baseline = {"value": v, "unit": "seq-steps/s", "cores": threads, "kind": "port",
                        "sample": f"{done} full learner iterations at batch {B} after 1 warm-up ({sec:.2f} s each), "
                                  f"{threads} torch threads (best of a probe over {os.cpu_count()} host cores)",
                        "cpu_model": cpu_model(), "host_cores": os.cpu_count()}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "seq-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (bf16x3 split tensor-core MMAs, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": f"{args.config}: " + " ".join(f"{k}={v}" for k, v in c.items()) + " per GPU",
                           "parallelism": f"dp{world}", "global_batch": world * B,
                           "replay_shard": f"{args.episodes} episodes x {ep_len + cfg.n_step} rows per GPU in HBM",
                           "l2": "inputs larger than L2: every step streams >1 GB of activations and gathers its batch "
                                 "from a multi-GB replay shard"},
                "e2e": {"value": e2e_value, "unit": "seq-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "feed": "pinned host batch -> staging buffer on a copy stream (step i+1 overlaps step i), "
                                "priorities + losses read back and the stream synchronised every step"},
                "iterations_per_s": world * 1e3 / ms, "rows_per_s": world * B * (cfg.burn_in + L) / (ms * 1e-3),
                "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "clocks": clk}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
