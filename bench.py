#!/usr/bin/env python
"""bench.py — OracleNet leaf evaluations per second at batch 1024 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B synthetic leaf positions: packed
bitboards (+ dM scalar, + legal-move CSR) -> planes -> OracleNet 6x128 -> softmax over the legal
moves and tanh(v_logit + k*dM).  Workload = BASELINE.json configs[1] at B = 1024 (the batch the
metric is quoted on).  Positions: seeded random legal playouts; weights: the reference's
initialisation law with live heads (oracle/weights.py variant B); both synthetic.

  value    whole-job evals/s with inputs already resident in HBM: K replays of the captured
           forward graph, each timed with CUDA events on the launching stream (inside the
           library), L2 flushed (256 MiB rewrite) between steps; max over ranks.
           value_sustained: the same forward replayed back to back for >= 2 s without the flush
           (the kernel is power-capped: the flush lets the clock recover, a busy pool does not).
  e2e      same metric through the public C ABI with HOST buffers: H2D of boards/q/CSR and D2H
           of priors/values inside the timed region, every step.  ONE host thread keeps two
           handles busy with cb_eval_leaves_submit / cb_eval_leaves_wait.
  roofline the dominant kernel — for 128-channel nets the ONE kernel of the forward (planes, 14 tcgen05
           convs, both heads) — algorithmic FLOPs per launch / mean CUDA-event duration of its launches,
           against the measured bf16 peak.
  cpu_baseline  the reference's own traced TorchScript module (oracle/_ref, built from the
           reference sources) on the host cores, bounded sample; "port" if _ref is absent.

oracle/ appears here in two roles only: as the generator of the synthetic inputs (seeded positions and
random-init weights, built before any timed region) and as the CPU baseline / --impl reference leg.  The
measured path is libcaissa_b200.so alone.

N > 1: one process per GPU (torchrun), each rank evaluates its own independent batch — the path
shards by position with no collective — so scaling is "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "neurosymbolic-mcts_b200"))

METRIC = "OracleNet leaf evals/sec @batch 1024"
UNIT = "evals/s"
FLOPS_PER_EVAL = {(6, 128): 249144320, (10, 256): 1593082880}
WEIGHT_SEED = 1803  # variant B seed of tests/golden/net_6x128_B.npz (value head alive)


def tower_flops_per_position(blocks, hidden):
    """Algorithmic FLOPs of the convolution tower (start conv with its real 17 input planes, 2 convs per
    block, p_conv); 2*MAC, padding not credited."""
    return 2 * 9 * 64 * (17 * hidden + (2 * blocks + 1) * hidden * hidden)


def workload_label(blocks, hidden, batch):
    """Which BASELINE.json config the arguments describe."""
    if (blocks, hidden) == (6, 128):
        return "configs[1]: batched leaf eval sweep, OracleNet 6x128, batch %d, one batch per GPU" % batch
    if (blocks, hidden) == (10, 256):
        return "configs[3]: scaled OracleNet 10x256 leaf eval, batch %d, one batch per GPU" % batch
    return "OracleNet %dx%d leaf eval, batch %d, one batch per GPU (not a BASELINE.json config)" % (blocks, hidden, batch)


def measured_traffic(blocks, hidden, batch, dtype):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of THIS configuration, or None."""
    tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if not os.path.exists(tp):
        return None
    rows = json.load(open(tp))
    key = "%dx%d_b%d_%s" % (blocks, hidden, batch, dtype)
    row = rows.get(key) if isinstance(rows, dict) else None
    return row.get("dram_bytes_per_launch") if row else None


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), "measured"
    return 1400.0, 1590.0, "fallback"  # B200_PROFILING.md fallback (sustained ~1.4, burst 1.59 PF)


def make_workload(args, rank):
    from oracle import pyoracle, weights as W
    boards, idx, off, raw = pyoracle.random_positions(20261019 + 7919 * rank, args.batch, 160)
    q = pyoracle.static_q(boards)
    blob = W.flatten(W.make_weights(args.blocks, args.hidden, WEIGHT_SEED, "B"), args.blocks, args.hidden)
    return boards, q, idx, off, blob, raw


class ClockSampler:
    """SM clock and throttle reasons sampled (NVML, every 5 ms) DURING the timed regions: the sampler thread
    records only while `active` is set, which the bench does around the resident and the end-to-end loops.
    (A 1 ms period perturbed the 8-GPU run: eight processes polling NVML cost one rank 40 % of its resident rate.)"""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], False, None
        self.max_mhz = None
        self.active = False

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))

            def run():
                while not self.stop:
                    if self.active:
                        try:
                            self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                              nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                        except Exception:
                            pass
                    time.sleep(0.005)
            self.t = threading.Thread(target=run, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t:
            self.t.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        bits = 0
        for _, r in self.rows:
            bits |= r
        return {"sm_mhz": float(np.median([c for c, _ in self.rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for n, b in names.items() if bits & b), "samples": len(self.rows)}


def cpu_reference_module(blocks, hidden):
    """The reference's traced TorchScript module if oracle/_ref has it, else the oracle port."""
    import torch
    name = "oraclenet_%dx%d_B.pt" % (blocks, hidden)
    path = os.path.join(REPO, "oracle", "_ref", name)
    if os.path.exists(path):
        return torch.jit.load(path, map_location="cpu").eval(), "reference"
    return None, "port"


def time_cpu_path(args, boards, q, idx, off, blob, budget_s, steps, warmup, moves_raw=None):
    """Reference CPU implementation of the path on a bounded sample: encode (C restatement of
    board_to_planes) -> traced OracleNet on libtorch CPU -> exp -> priors (C restatement)."""
    import torch
    from oracle import net, pyoracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    module, kind = cpu_reference_module(args.blocks, args.hidden)
    raw = [pyoracle.array_to_board(b) for b in boards]

    def one_step(n):
        planes = pyoracle.planes_batch(boards[:n])
        if module is not None:
            scal = torch.from_numpy(np.stack([q[:n], np.ones(n, dtype=np.float32)], axis=1))
            with torch.no_grad():
                logp, v, k = module(torch.from_numpy(planes), scal)
            probs = logp.exp().numpy()
        else:
            logp, v, k = net.forward(blob, args.blocks, args.hidden, planes, q[:n])
            probs = np.exp(logp)
        for i in range(n):   # policy_to_move_priors: the C restatement (src/tensor.rs:184-213)
            pyoracle.policy_to_move_priors(probs[i], moves_raw[off[i]:off[i + 1]], raw[i].w_to_move)
        return n

    n = min(256, args.batch)
    one_step(min(n, 32))
    t0 = time.perf_counter(); one_step(n); t1 = time.perf_counter() - t0
    # size the per-step sample so the whole (warmup + steps) run fits the budget
    per_pos = t1 / n
    n = int(max(16, min(args.batch, budget_s / max(steps + warmup, 1) / per_pos)))
    for _ in range(warmup):
        one_step(n)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step(n)
    dt = time.perf_counter() - t0
    del raw
    return {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d steps x %d positions of the same batch (encode + traced OracleNet fp32 on libtorch CPU "
                      "+ exp + legal-move priors), %d torch threads" % (steps, n, cores)}, dt / steps * 1e3


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else written to fd 1 while the
    benchmark runs (NCCL's version banner, library chatter) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def resident_rate(ev, boards, q, idx, off, B, iters, warmup=3):
    """Flushed-L2 resident rate of one batch size: median CUDA-event time of `iters` replays -> (evals/s, us)."""
    ev.stage_leaves(boards[:B], q[:B], idx[:off[B]], off[:B + 1])
    ev.time_resident(warmup, material_on=True, flush_l2=True)
    ms = ev.time_resident(iters, material_on=True, flush_l2=True)
    med = float(np.median(ms))
    return B / (med * 1e-3), med * 1e3


def config4_record(cb, local_rank, peaks):
    """BASELINE.json configs[3] on this GPU: OracleNet 10x256 bf16, batch 2048 (tower_r2_umma_kernel on CTA pairs)."""
    from oracle import pyoracle, weights as W
    blocks, hidden, B, dtype = 10, 256, 2048, "bf16"
    boards, idx, off, _ = pyoracle.random_positions(4242, B, 160)
    q = pyoracle.static_q(boards)
    blob = W.flatten(W.make_weights(blocks, hidden, WEIGHT_SEED, "B"), blocks, hidden)
    ev = cb.Evaluator(blocks, hidden, dtype, device=local_rank)
    ev.load_weights(blob)
    rate, us = resident_rate(ev, boards, q, idx, off, B, 40)
    ev.time_tower(2)
    tower_ms, n_tower = ev.time_tower(16)
    sus_ms, sus_n = ev.time_sustained(1.0)
    launches = ev.launches_per_forward()
    # config 4's other half: self-play with this net (1024 concurrent games, 200 sims/move, tree operations on the device)
    sp = ev.selfplay(games=1024, sims_per_move=200, material_on=False, seed=77, max_seconds=1.5, pipeline=3)
    ev.close()
    sustained, burst, _ = peaks
    tower_tf = tower_flops_per_position(blocks, hidden) * B / (float(tower_ms.mean()) * 1e-3) / 1e12
    return {"workload": workload_label(blocks, hidden, B), "dtype": dtype, "value": rate, "unit": UNIT, "us_per_forward": us,
            "timing": "value: median of 40 flushed steps (CUDA events); roofline: mean of 16 tower launches; value_sustained: "
                      "back-to-back replays for 1 s.  Short regions run at up to 1.965 GHz, long ones at ~1.79 GHz under the "
                      "power cap: `python bench.py --blocks 10 --hidden 256 --batch 2048` gives the 200-step figures",
            "value_sustained": B * sus_n / (sus_ms * 1e-3), "gpu_launches_per_forward": launches,
            "selfplay_sims_per_s": sp["sims"] / sp["seconds"], "selfplay_games": 1024,
            "roofline": {"bound": "tensor", "kernel": "tower_r2_umma_kernel<bf16> (planes, conv tower, policy head on CTA pairs, "
                                                    "cta_group::2; %d launch/step)" % n_tower,
                         "achieved": tower_tf, "peak": burst, "unit": "TFLOP/s", "frac": tower_tf / burst,
                         "traffic": measured_traffic(blocks, hidden, B, dtype)},
            "parity": "tests/test_gpu_parity.py::test_tensor_core_modes_match_reference[10x256_B-bf16]: |dV| <= 1e-2, KL <= 1e-4"}


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the 10x256 batch-2048 sub-record (N = 1 only)")
    ap.add_argument("--selfplay-seconds", type=float, default=3.0, help="seconds per self-play pool shape; 0 disables the section")
    ap.add_argument("--selfplay-games", type=int, default=1024)
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the back-to-back replay loop")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": workload_label(args.blocks, args.hidden, args.batch), "batch_per_gpu": args.batch,
              "blocks": args.blocks, "hidden": args.hidden,
              "positions": "seeded random legal playouts (oracle chess substrate), material dM = static PeSTO/100",
              "weights": "reference init law, heads de-zeroed (variant B, numpy seed %d)" % WEIGHT_SEED,
              "l2": "flushed (256 MiB rewrite) between timed steps", "material_on": True}

    if args.impl == "reference":
        if rank != 0:
            return 0
        boards, q, idx, off, blob, raw = make_workload(args, 0)
        cpu, ms = time_cpu_path(args, boards, q, idx, off, blob, 150.0, args.steps, args.warmup, raw)
        line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    import caissa_b200 as cb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the only NCCL traffic is the off-path exchange: keep its kernels to a few SMs (the pools need the rest)
        os.environ.setdefault("NCCL_MAX_CTAS", "4")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    boards, q, idx, off, blob, raw = make_workload(args, rank)
    ev = cb.Evaluator(args.blocks, args.hidden, args.dtype, device=local_rank)
    ev.load_weights(blob)
    B, K, W = args.batch, args.steps, args.warmup
    peaks = load_peaks()

    # ---- device-resident throughput (value): exactly K timed steps
    ev.stage_leaves(boards, q, idx, off)
    ev.time_resident(W, material_on=True, flush_l2=True)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    clocks.active = True
    t_wall0 = time.perf_counter()
    ms_each = ev.time_resident(K, material_on=True, flush_l2=True)
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks.active = False
    dev_ms = float(ms_each.sum())
    pri, v, vf, kk = ev.read_resident()
    assert np.isfinite(pri).all() and np.isfinite(vf).all() and np.abs(vf).max() <= 1.0

    # ---- roofline of the dominant kernel: the fused persistent conv tower (one launch per step), timed alone like `value`
    ev.time_tower(3)
    tower_ms, n_tower = ev.time_tower(max(K // 2, 5))
    launch_ms = float(tower_ms.mean())
    single_launch = ev.launches_per_forward() == 1     # resident tower: planes, convs and both heads in one kernel
    if single_launch:
        flops_launch = FLOPS_PER_EVAL[(args.blocks, args.hidden)] * B
    else:
        flops_launch = tower_flops_per_position(args.blocks, args.hidden) * B

    # ---- sustained: the same forward back to back for >= 2 s (no flush, no pause: the power-capped rate)
    barrier()
    clocks.active = True
    sus_ms, sus_n = ev.time_sustained(args.sustained_seconds)
    clocks.active = False
    barrier()
    time.sleep(0.5)                                    # let the board cool before the next timed region

    # ---- end to end through the public C ABI with host buffers: every step copies boards / q / CSR host->device and
    # priors / values device->host inside the timed region.  ONE host thread drives two evaluator handles (independent
    # streams, the ABI's documented multi-handle use) with cb_eval_leaves_submit / _wait, so one handle's copies run
    # under the other's kernel; the single-handle synchronous rate is kept too.  (Host buffers and argument pointers
    # are set up once, as a host that reuses its request buffers would.)
    run1 = ev.prepared_eval_leaves(boards, q, idx, off)
    for _ in range(W):
        run1()
    barrier()
    clocks.active = True
    t0 = time.perf_counter()
    for _ in range(K):
        out = run1()
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    clocks.active = False
    ev2 = cb.Evaluator(args.blocks, args.hidden, args.dtype, device=local_rank)
    ev2.load_weights(blob)
    sub = [ev.prepared_submit_wait(boards, q, idx, off), ev2.prepared_submit_wait(boards, q, idx, off)]

    def pipelined(n):
        """n batches through the two handles from this one thread, at most one batch in flight per handle."""
        sub[0][0]()
        for i in range(1, n):
            sub[i & 1][0]()
            sub[(i - 1) & 1][1]()
        return sub[(n - 1) & 1][1]()
    pipelined(max(W, 4))
    barrier()
    clocks.active = True
    t0 = time.perf_counter()
    out = pipelined(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the same loop for >= 0.5 s whatever --steps says (a short K leaves a timed region of a few ms)
    n_long = max(K, int(0.5 / max(e2e_s / K, 1e-6)))
    t0 = time.perf_counter()
    pipelined(n_long)
    torch.cuda.synchronize()
    e2e_long_s = time.perf_counter() - t0
    clocks.active = False
    in_region = len(clocks.rows)
    clocks.__exit__()
    clk = clocks.summary()
    clk["regions"] = ("NVML every 5 ms during the resident loop, the sustained loop and the end-to-end loops (the timed "
                      "regions): %d samples" % in_region)
    barrier()
    h2d = int(boards.nbytes + q.nbytes + idx.nbytes + off.nbytes)
    d2h = int(out[0].nbytes + out[1].nbytes + out[2].nbytes + out[3].nbytes)

    # ---- batch sweep (config 2's small end: the reference's server batches <= 16 by default) and config 4, rank 0 of N = 1
    sweep = cfg4 = None
    if world == 1:
        sweep = {}
        for b_ in (1, 16, 128, 256):
            if b_ <= B:
                rate, us = resident_rate(ev, boards, q, idx, off, b_, 30)
                sweep[str(b_)] = {"evals_per_s": rate, "us_per_forward": us}
        ev.stage_leaves(boards, q, idx, off)
        # the other arithmetic modes of the same net and batch: fp16 operands (same kernel), and the mode that meets the
        # north star's fp32 tolerance: fp16-PAIR operands on the tensor cores (three MMAs per product) for 128-channel nets,
        # fp32 FMA on CUDA cores otherwise
        modes = {}
        for dt in ("fp16", "fp32"):
            if dt != args.dtype:
                evm = cb.Evaluator(args.blocks, args.hidden, dt, device=local_rank)
                evm.load_weights(blob)
                rate, us = resident_rate(evm, boards, q, idx, off, B, 10, warmup=2)
                split = dt == "fp32" and evm.launches_per_forward() == 1
                modes[dt] = {"evals_per_s": rate, "us_per_forward": us,
                             "arithmetic": "fp16 operands, fp32 accumulate" if dt == "fp16" else
                                           ("fp16 pairs (x = hi + lo), W_hi X_hi + W_hi X_lo + W_lo X_hi on tcgen05, fp32 accumulate"
                                            if split else "fp32 FMA on CUDA cores"),
                             "parity_bar": "|dV| <= 1e-2, KL <= 1e-4" if dt == "fp16" else
                                           "|dV| <= 1e-3, KL <= 1e-4 (measured %s)" % ("1.2e-5 / 5e-7" if split else "4e-7 / 2e-6")}
                evm.close()
        sweep["modes_at_batch_%d" % B] = modes
        if not args.no_config4 and (args.blocks, args.hidden) == (6, 128):
            cfg4 = config4_record(cb, local_rank, peaks)

    # ---- self-play pool on this GPU (config 3 shape: 200 sims/move, 1024 concurrent games, vanilla mode): the
    # reference's search (tactical-first phase, f64 PUCT, tree reuse) with the tree operations on the device, and beside
    # it the host-driven pool (two half-pools, each with its own evaluator handle).
    sp = sp_host = sp_koth = sp_two = sp_two_pe = sp_rec = sp_deep = sp_deep4 = None
    nthreads = max(1, (os.cpu_count() or 1) // world - 2)
    rec_path = "/dev/shm/caissa_bench_samples_r%d.bin" % rank
    if args.selfplay_seconds > 0:
        T = args.selfplay_seconds
        barrier()
        sp = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=1 + rank, max_seconds=T, pipeline=3)
        barrier()
        sp_host = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=101 + rank,
                              max_seconds=T, threads=nthreads, pipeline=2, aux=ev2)
        barrier()
        # config 5 shape: King-of-the-Hill rules, 800 simulations per move
        sp_koth = ev.selfplay(games=args.selfplay_games, sims_per_move=800, material_on=False, koth=True, seed=201 + rank,
                              max_seconds=T, pipeline=3)
        barrier()
        # two pools of that many games each: one pool's tree kernel runs beside the other pool's forward kernel;
        # vanilla, and with the material term from the principal-exchange q-search run on the device for every leaf
        sp_two = ev.selfplay(games=2 * args.selfplay_games, sims_per_move=200, material_on=False, seed=301 + rank,
                             max_seconds=T, pipeline=4)
        barrier()
        sp_two_pe = ev.selfplay(games=2 * args.selfplay_games, sims_per_move=200, material_on=True, qsearch="pe",
                                seed=401 + rank, max_seconds=T, pipeline=4)
        barrier()
        # the reference's own self-play configuration of Tier 1 (src/bin/self_play.rs:213,241-283): at every leaf the mate search
        # to mate-in-5 by checking moves, at every new root mate_search(5, exhaustive 2); dM by the exchange line.  The gate
        # searches run on the device as resumable jobs (a game whose job is suspended completes no simulation in that step, so
        # the rate depends on the games in flight: the configured pool and one four times as large)
        sp_deep = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=True, qsearch="pe", tier1_light="deep",
                              seed=601 + rank, max_seconds=T, pipeline=3)
        barrier()
        sp_deep4 = ev.selfplay(games=4 * args.selfplay_games, sims_per_move=200, material_on=True, qsearch="pe", tier1_light="deep",
                               seed=701 + rank, max_seconds=T, pipeline=3)
        barrier()
        # the one-pool run again with the training samples produced: every finished game is assembled on the host into
        # the reference's 5763-f32 records and written to a file (tmpfs) — the records the NCCL gather below moves
        if os.path.exists(rec_path):
            os.remove(rec_path)
        sp_rec = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=501 + rank,
                             max_seconds=T, pipeline=3, sample_path=rec_path)
        barrier()

    # ---- off the hot path: the exchanges the multi-GPU design has (NCCL over NVLink), with REAL data: the sample records
    # this rank's self-play just wrote are gathered (all ranks -> every rank), checked bit for bit on rank 0, then gathered
    # again on a side thread WHILE a self-play pool runs (the cost of the exchange to the hot path); a new weight blob is
    # broadcast from rank 0, loaded by every rank's evaluator and spot-checked against rank 0's outputs.
    nccl = None
    if world > 1 and sp_rec is not None:
        import zlib
        from caissa_b200 import distributed as D
        local = np.fromfile(rec_path, dtype=np.float32).reshape(-1, D.RECORD_FLOATS) if os.path.exists(rec_path) else \
            np.zeros((0, D.RECORD_FLOATS), dtype=np.float32)
        ns = [None] * world
        dist.all_gather_object(ns, int(local.shape[0]))
        # the records travel in their lossless sparse wire form (caissa_b200/distributed.py pack_records: ~0.9 KB instead
        # of 23,052 B per record); rank 0 expands them back into the reference's dense records
        D.gather_samples_packed(local[:min(8, local.shape[0])], dst=0)   # communicator and kernels warm
        torch.cuda.synchronize()
        dist.barrier()
        gt = {}
        t0 = time.perf_counter()
        allrec, wire_bytes = D.gather_samples_packed(local, dst=0, timings=gt)
        torch.cuda.synchronize()
        gather_s = time.perf_counter() - t0
        wires = [None] * world
        dist.all_gather_object(wires, (int(wire_bytes), gt.get("pack_s", 0.0), gt.get("wire_s", 0.0), gt.get("unpack_s", 0.0)))
        sums = [None] * world
        dist.all_gather_object(sums, (int(local.shape[0]), zlib.crc32(local.tobytes())))
        ok = True
        if rank == 0:
            lo = 0
            for n_r, crc in sums:
                ok = ok and zlib.crc32(allrec[lo:lo + n_r].tobytes()) == crc
                lo += n_r
            ok = ok and lo == allrec.shape[0]
        # the gather under load: a side thread gathers this rank's records in a FIXED number of flushes (the same on every
        # rank: collectives must pair up) spread over the run of a self-play pool
        # Small frequent flushes: the cost to the pool grows with the length of each NCCL kernel, not with the bytes per
        # second (tools/exchange_probe.py, 4 x B200: 100 x 128 records/s costs 0.6 %, 25 x 512 costs 2.9 %, 5 x 2560 more).
        n_flush = max(1, int(args.selfplay_seconds * 50))
        flushes = [0]
        host_group = dist.new_group(backend="gloo")   # block sizes are agreed on the host: ONE NCCL kernel per flush, ranks aligned

        def gather_loop():
            torch.cuda.set_device(local_rank)
            chunk = local[:max(1, min(local.shape[0], 256))]
            for _ in range(n_flush):          # 50 flushes/s x 256 records/rank (12.8 k records/s/rank): above the produced rate
                D.gather_samples_packed(chunk, dst=0, size_group=host_group)
                flushes[0] += 1
                time.sleep(0.012)
        dist.barrier()
        th = threading.Thread(target=gather_loop)
        th.start()
        # (two pools: the forward kernel of a pool runs on 128 SMs, so the tree kernel of the other pool AND the NCCL kernels
        #  find free SMs.  A one-pool forward is one wave of 148 single-SM CTAs: any other kernel holding an SM — an NCCL
        #  kernel waiting for its peer does for hundreds of microseconds — pushes one CTA into a second wave and doubles the
        #  step, which at 8 ranks cost 43 % of the sims/s when measured that way.)
        sp_g = ev.selfplay(games=2 * args.selfplay_games, sims_per_move=200, material_on=False, seed=301 + rank,
                           max_seconds=args.selfplay_seconds, pipeline=4)
        th.join()
        # weight refresh: rank 0 draws a new blob, NCCL broadcast, every rank loads it and evaluates the same 8 positions
        from oracle import weights as Wt
        new_blob = Wt.flatten(Wt.make_weights(args.blocks, args.hidden, WEIGHT_SEED + 1, "B"), args.blocks, args.hidden) \
            if rank == 0 else None
        t0 = time.perf_counter()
        got = D.broadcast_weights(new_blob, ev.weight_count())
        bcast_s = time.perf_counter() - t0
        ev2.load_weights(got)
        b0, q0, i0, o0, _, _ = make_workload(argparse.Namespace(batch=8, blocks=args.blocks, hidden=args.hidden), 0)
        _, v0, _, _ = ev2.eval_leaves(b0, q0, i0, o0)
        vs = [None] * world
        dist.all_gather_object(vs, v0.tobytes())
        n_all = int(sum(n_r for n_r, _ in sums))
        nccl = {"gather_records": n_all, "gather_seconds": gather_s,
                "gather_GBps": n_all * D.RECORD_FLOATS * 4 / gather_s / 1e9, "gather_bit_equal_on_rank0": bool(ok),
                "gather": "every rank's finished-game records of the run above -> rank 0 in the lossless sparse wire form "
                          "(pack on every rank, one NCCL all-gather, host copy and expansion to the dense 5763-f32 records on "
                          "rank 0 only); gather_GBps counts the dense bytes delivered",
                "gather_wire_bytes": int(sum(w[0] for w in wires)), "gather_dense_bytes": n_all * D.RECORD_FLOATS * 4,
                "gather_pack_seconds_max": max(w[1] for w in wires), "gather_wire_seconds_rank0": wires[0][2],
                "gather_unpack_seconds_rank0": wires[0][3],
                "records_per_s_produced": None, "selfplay_seconds_with_gather": sp_g["seconds"], "selfplay_sims_with_gather": sp_g["sims"],
                "gather_flushes_during_selfplay": int(flushes[0]),
                "weights_broadcast_floats": int(got.size), "weights_broadcast_seconds": bcast_s,
                "weights_spot_check_identical_on_all_ranks": all(x == vs[0] for x in vs)}
    ev2.close()

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3, launch_ms, e2e_sync_s * 1e3, sp["seconds"] if sp else 0.0,
                      sp_host["seconds"] if sp_host else 0.0, sp_koth["seconds"] if sp_koth else 0.0,
                      sp_two["seconds"] if sp_two else 0.0, sp_two_pe["seconds"] if sp_two_pe else 0.0,
                      sp_rec["seconds"] if sp_rec else 0.0, sus_ms / max(sus_n, 1), e2e_long_s * 1e3 / n_long,
                      nccl["selfplay_seconds_with_gather"] if nccl else 0.0,
                      sp_deep["seconds"] if sp_deep else 0.0, sp_deep4["seconds"] if sp_deep4 else 0.0],
                     dtype=torch.float64, device="cuda")
    tot = torch.tensor([sp["sims"] if sp else 0, sp["nn_evals"] if sp else 0, sp_host["sims"] if sp_host else 0,
                        sp_koth["sims"] if sp_koth else 0, sp_two["sims"] if sp_two else 0,
                        sp_two_pe["sims"] if sp_two_pe else 0, sp_rec["sims"] if sp_rec else 0,
                        sp_rec["samples"] if sp_rec else 0, nccl["selfplay_sims_with_gather"] if nccl else 0,
                        sum(x["overflow"] for x in (sp, sp_koth, sp_two, sp_two_pe, sp_rec, sp_deep, sp_deep4) if x),
                        sp_deep["sims"] if sp_deep else 0, sp_deep4["sims"] if sp_deep4 else 0,
                        sp_deep["gate_hits"] if sp_deep else 0, sp_deep["gate_waits"] if sp_deep else 0,
                        sp_deep["steps"] if sp_deep else 0],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    (dev_ms, e2e_ms, launch_ms, e2e_sync_ms, sp_seconds, sp_host_seconds, sp_koth_seconds, sp_two_seconds,
     sp_two_pe_seconds, sp_rec_seconds, sus_ms_per_launch, e2e_long_ms_per_step, sp_g_seconds, sp_deep_seconds,
     sp_deep4_seconds) = [float(x) for x in t.tolist()]
    value = world * B * K / (dev_ms * 1e-3)
    e2e_value = world * B * K / (e2e_ms * 1e-3)

    if rank == 0:
        sustained, burst, how = peaks
        achieved = flops_launch / (launch_ms * 1e-3) / 1e12
        value_sustained = world * B / (sus_ms_per_launch * 1e-3)
        sus_tf = value_sustained / world * FLOPS_PER_EVAL.get((args.blocks, args.hidden), 0) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "value_sustained": value_sustained,
            "sustained": {"seconds": args.sustained_seconds, "launches": sus_n, "us_per_forward": sus_ms_per_launch * 1e3,
                          "whole_path_tflops": sus_tf, "peak": sustained, "frac": sus_tf / sustained,
                          "note": "the staged forward replayed back to back, no L2 flush, no pause: the power-capped rate; "
                                  "denominator = bf16_tflops_sustained of MEASURED_PEAKS.json"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K,
                    "api": "cb_eval_leaves_submit / cb_eval_leaves_wait with host buffers, 2 handles driven by ONE host thread per GPU",
                    "sustained_value": world * B / (e2e_long_ms_per_step * 1e-3), "sustained_steps": n_long,
                    "single_handle_sync_value": world * B * K / (e2e_sync_ms * 1e-3)},
            "gpu_launches": int(ev.launches_per_forward() * K * world),
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": achieved / burst, "traffic": measured_traffic(args.blocks, args.hidden, B, args.dtype),
                         "kernel": ("tower_r_umma_kernel<%s> (whole forward: planes, conv tower, policy and value heads; "
                                    "%d launch/step)" % (args.dtype, n_tower)) if single_launch else
                                   ("tower_r2_umma_kernel<%s> (planes, conv tower, policy head on CTA pairs; %d launch/step)" % (
                                       args.dtype, n_tower) if args.hidden == 256 else
                                    "tower_umma_kernel<%d> (whole conv tower, %d launch/step)" % (args.hidden, n_tower)),
                         "flops_per_launch": flops_launch, "launch_ms": launch_ms,
                         "peak_source": "bf16_tflops (burst: the tower launch is timed alone with CUDA events) of "
                                        "MEASURED_PEAKS.json (%s); sustained figure %.1f -> frac %.3f" % (
                                            how, sustained, achieved / sustained)},
            "whole_path_tflops": value / world * FLOPS_PER_EVAL.get((args.blocks, args.hidden), 0) / 1e12,
            "wall_s_timed_region": wall,
        }
        if sweep:
            line["batch_sweep"] = sweep
        if cfg4:
            line["config4"] = cfg4
        if sp:
            line["selfplay"] = {"sims_per_s": float(tot[0].item()) / sp_seconds,
                                "nn_evals_per_s": float(tot[1].item()) / sp_seconds,
                                "games_per_gpu": args.selfplay_games, "sims_per_move": 200,
                                "mode": "vanilla (config 5 rules: Tier 1 off, material off), one leaf in flight per game; the "
                                        "reference's search — tactical-first phase, root force-visit, f64 PUCT, tree reuse — with "
                                        "the tree operations on the device (cb_selfplay_run pipeline 3: select, forward, "
                                        "expand/play kernels in one CUDA graph); trees bit-equal to oracle/mcts.c (tests)",
                                "seconds": sp_seconds, "mean_batch": sp["mean_batch"],
                                "forward_fraction": sp["eval_seconds"] / max(sp["seconds"], 1e-9),
                                "host_driven_sims_per_s": float(tot[2].item()) / max(sp_host_seconds, 1e-9),
                                "koth_800_sims_per_move_sims_per_s": float(tot[3].item()) / max(sp_koth_seconds, 1e-9),
                                "two_pools_sims_per_s": float(tot[4].item()) / max(sp_two_seconds, 1e-9),
                                "two_pools_material_pe_sims_per_s": float(tot[5].item()) / max(sp_two_pe_seconds, 1e-9),
                                "with_sample_records_sims_per_s": float(tot[6].item()) / max(sp_rec_seconds, 1e-9),
                                "sample_records_per_s": float(tot[7].item()) / max(sp_rec_seconds, 1e-9),
                                "arena_overflows": int(tot[9].item()),
                                "two_pools_games_per_gpu": 2 * args.selfplay_games,
                                "two_pools_mode": "cb_selfplay_run pipeline 4: two pools of games_per_gpu games alternate, "
                                                  "the tree kernel of one beside the forward kernel of the other; "
                                                  "material_pe: dM by the principal-exchange line on the device",
                                "host_threads_per_gpu": nthreads}
            if sp_deep:
                line["selfplay"]["tier1_deep"] = {
                    "mode": "the reference's Tier-1 gate as its self-play runs it (tier1_light = 2): leaf gate mate_search to "
                            "mate-in-5 by checking moves, every new root adjudicated by mate_search(5, exhaustive 2); dM by the "
                            "exchange line; the gate searches are resumable any/all minimax jobs on the device (mcts_gate_kernel), "
                            "trees bit-equal to oracle/mcts.c + oracle/gates.c (tests).  Measured from the start position for "
                            "the run's seconds; the steady-state figures of longer runs are in profiles/r02_deep_gates_probe_long.json",
                    "sims_per_s": float(tot[10].item()) / max(sp_deep_seconds, 1e-9), "games_per_gpu": args.selfplay_games,
                    "sims_per_s_4x_games": float(tot[11].item()) / max(sp_deep4_seconds, 1e-9), "games_per_gpu_4x": 4 * args.selfplay_games,
                    "leaves_decided_by_the_gate": int(tot[12].item()),
                    "game_steps_waiting_for_a_gate_job_fraction": float(tot[13].item()) / max(float(tot[14].item()) * args.selfplay_games, 1.0)}
        if nccl:
            nccl["records_per_s_produced"] = float(tot[7].item()) / max(sp_rec_seconds, 1e-9)
            nccl["sims_per_s_with_gather_running"] = float(tot[8].item()) / max(sp_g_seconds, 1e-9)
            nccl["sims_per_s_without"] = float(tot[4].item()) / max(sp_two_seconds, 1e-9)
            nccl["pool"] = "two pools of %d games (pipeline 4)" % args.selfplay_games
            for k_ in ("selfplay_seconds_with_gather", "selfplay_sims_with_gather"):
                nccl.pop(k_)
            line["nccl_exchange"] = nccl
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = time_cpu_path(args, boards, q, idx, off, blob, 20.0, 3, 1, raw)
            line["cpu_baseline"] = cpu
        emit(line)
    if os.path.exists(rec_path):
        os.remove(rec_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
