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
  e2e      same metric through the public C-ABI call cb_eval_leaves with HOST buffers: H2D of
           boards/q/CSR and D2H of priors/values inside the timed region, every step.
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


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), "measured"
    return 1400.0, 1590.0, "fallback"  # B200_PROFILING.md fallback (sustained ~1.4, burst 1.59 PF)


def make_workload(args, rank):
    from oracle import pyoracle, weights as W
    boards, idx, off, _ = pyoracle.random_positions(20261019 + 7919 * rank, args.batch, 160)
    q = pyoracle.static_q(boards)
    blob = W.flatten(W.make_weights(args.blocks, args.hidden, WEIGHT_SEED, "B"), args.blocks, args.hidden)
    return boards, q, idx, off, blob


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


def time_cpu_path(args, boards, q, idx, off, blob, budget_s, steps, warmup):
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
        for i in range(n):
            m = idx[off[i]:off[i + 1]]
            p = probs[i][m]
            s = p.sum()
            p = p / s if s > 0 else np.full(len(m), 1.0 / max(len(m), 1), dtype=np.float32)
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
    ap.add_argument("--selfplay-seconds", type=float, default=3.0, help="seconds per self-play pool shape (two are run); 0 disables the section")
    ap.add_argument("--selfplay-games", type=int, default=1024)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "configs[1]: batched leaf eval, OracleNet %dx%d, batch %d, one batch per GPU" % (
        args.blocks, args.hidden, args.batch)
    config = {"workload": workload, "batch_per_gpu": args.batch, "blocks": args.blocks, "hidden": args.hidden,
              "positions": "seeded random legal playouts (oracle chess substrate), material dM = static PeSTO/100",
              "weights": "reference init law, heads de-zeroed (variant B, numpy seed %d)" % WEIGHT_SEED,
              "l2": "flushed (256 MiB rewrite) between timed steps", "material_on": True}

    if args.impl == "reference":
        if rank != 0:
            return 0
        boards, q, idx, off, blob = make_workload(args, 0)
        cpu, ms = time_cpu_path(args, boards, q, idx, off, blob, 150.0, args.steps, args.warmup)
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    boards, q, idx, off, blob = make_workload(args, rank)
    ev = cb.Evaluator(args.blocks, args.hidden, args.dtype, device=local_rank)
    ev.load_weights(blob)
    B, K, W = args.batch, args.steps, args.warmup

    # ---- device-resident throughput (value)
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

    # ---- roofline of the dominant kernel: the fused persistent conv tower (one launch per step)
    ev.time_tower(3)
    tower_ms, n_tower = ev.time_tower(max(K // 2, 5))
    launch_ms = float(tower_ms.mean())
    single_launch = ev.launches_per_forward() == 1     # resident tower: planes, convs and both heads in one kernel
    if single_launch:
        flops_launch = FLOPS_PER_EVAL[(args.blocks, args.hidden)] * B
    else:
        flops_launch = tower_flops_per_position(args.blocks, args.hidden) * B

    # ---- end to end through the public C ABI with host buffers: every step copies boards / q / CSR
    # host->device and priors / values device->host inside the timed region.  Two evaluator handles
    # (independent streams, the ABI's documented multi-handle use) are driven by two host threads so
    # one handle's copies overlap the other's kernels; the single-handle synchronous rate is kept too.
    # (host buffers and argument pointers are set up once, as a host that reuses its request buffers would;
    #  every call still copies inputs H2D and results D2H inside cb_eval_leaves)
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
    run2 = ev2.prepared_eval_leaves(boards, q, idx, off)
    for _ in range(W):
        run2()
    counts = [K - K // 2, K // 2]

    def pump(e, n):
        run = run1 if e is ev else run2
        for _ in range(n):
            run()
    barrier()
    threads = [threading.Thread(target=pump, args=(e, n)) for e, n in zip((ev, ev2), counts)]
    clocks.active = True
    t0 = time.perf_counter()
    [t.start() for t in threads]
    [t.join() for t in threads]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks.active = False
    in_region = len(clocks.rows)
    if in_region < 5:
        # short timed regions (few steps): top the sample up under the same load, outside the timing
        clocks.active = True
        ev.time_resident(max(200, K), material_on=True, flush_l2=True)
        clocks.active = False
    clocks.__exit__()
    clk = clocks.summary()
    clk["regions"] = ("NVML every 5 ms during the resident loop and both end-to-end loops (the timed regions): %d samples; "
                      "%d more under an untimed replay of the resident loop" % (in_region, len(clocks.rows) - in_region))
    barrier()
    h2d = int(boards.nbytes + q.nbytes + idx.nbytes + off.nbytes)
    d2h = int(out[0].nbytes + out[1].nbytes + out[2].nbytes + out[3].nbytes)

    # ---- self-play pool on this GPU (config 3 shape: 200 sims/move, 1024 concurrent games, vanilla mode):
    # tree operations on the device (select / expand / backup / move choice as two kernels around the forward
    # kernel, no host work per simulation), and beside it the host-driven pool (two half-pools of 512 games,
    # each with its own evaluator handle) whose rate depends on the host cores available per GPU.
    sp = sp_host = sp_koth = sp_two = sp_two_pe = sp_rec = None
    nthreads = max(1, (os.cpu_count() or 1) // world - 2)
    if args.selfplay_seconds > 0:
        barrier()
        sp = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=1 + rank,
                         max_seconds=args.selfplay_seconds, pipeline=3)
        barrier()
        sp_host = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=101 + rank,
                              max_seconds=args.selfplay_seconds, threads=nthreads, pipeline=2, aux=ev2)
        barrier()
        # config 5 shape: King-of-the-Hill rules, 800 simulations per move
        sp_koth = ev.selfplay(games=args.selfplay_games, sims_per_move=800, material_on=False, koth=True, seed=201 + rank,
                              max_seconds=args.selfplay_seconds, pipeline=3)
        barrier()
        # two pools of that many games each: one pool's tree kernel runs beside the other pool's forward kernel;
        # vanilla, and with the material term from the principal-exchange q-search run on the device for every leaf
        sp_two = ev.selfplay(games=2 * args.selfplay_games, sims_per_move=200, material_on=False, seed=301 + rank,
                             max_seconds=args.selfplay_seconds, pipeline=4)
        barrier()
        sp_two_pe = ev.selfplay(games=2 * args.selfplay_games, sims_per_move=200, material_on=True, qsearch="pe",
                                seed=401 + rank, max_seconds=args.selfplay_seconds, pipeline=4)
        barrier()
        # the one-pool run again with the training samples produced: every finished game is assembled on the host into
        # the reference's 5763-f32 records and written out (to /dev/null: record building and write calls, no storage)
        sp_rec = ev.selfplay(games=args.selfplay_games, sims_per_move=200, material_on=False, seed=501 + rank,
                             max_seconds=args.selfplay_seconds, pipeline=3, sample_path="/dev/null")
        barrier()
    ev2.close()

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3, launch_ms, e2e_sync_s * 1e3, sp["seconds"] if sp else 0.0,
                      sp_host["seconds"] if sp_host else 0.0, sp_koth["seconds"] if sp_koth else 0.0,
                      sp_two["seconds"] if sp_two else 0.0, sp_two_pe["seconds"] if sp_two_pe else 0.0,
                      sp_rec["seconds"] if sp_rec else 0.0],
                     dtype=torch.float64, device="cuda")
    tot = torch.tensor([sp["sims"] if sp else 0, sp["nn_evals"] if sp else 0, sp_host["sims"] if sp_host else 0,
                        sp_koth["sims"] if sp_koth else 0, sp_two["sims"] if sp_two else 0,
                        sp_two_pe["sims"] if sp_two_pe else 0, sp_rec["sims"] if sp_rec else 0,
                        sp_rec["samples"] if sp_rec else 0],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    (dev_ms, e2e_ms, launch_ms, e2e_sync_ms, sp_seconds, sp_host_seconds, sp_koth_seconds, sp_two_seconds,
     sp_two_pe_seconds, sp_rec_seconds) = [float(x) for x in t.tolist()]
    value = world * B * K / (dev_ms * 1e-3)
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    # off the hot path: the one exchange the multi-GPU design has (training-sample gather over NCCL)
    gathered = None
    if world > 1:
        from caissa_b200 import distributed as D
        rec = np.zeros((4, D.RECORD_FLOATS), dtype=np.float32) + rank
        gathered = int(D.gather_samples(rec).shape[0])

    if rank == 0:
        sustained, burst, how = load_peaks()
        achieved = flops_launch / (launch_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K,
                    "api": "cb_eval_leaves with host buffers, 2 handles x 2 host threads per GPU",
                    "single_handle_sync_value": world * B * K / (e2e_sync_ms * 1e-3)},
            "gpu_launches": int(ev.launches_per_forward() * K * world),
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": achieved / burst, "traffic": traffic,
                         "kernel": ("tower_r_umma_kernel<%s> (whole forward: planes, conv tower, policy and value heads; "
                                    "%d launch/step)" % (args.dtype, n_tower)) if single_launch else
                                   ("tower_umma_kernel<%d> (whole conv tower, %d launch/step)" % (args.hidden, n_tower)),
                         "flops_per_launch": flops_launch, "launch_ms": launch_ms,
                         "peak_source": "bf16_tflops (burst: the tower launch is timed alone with CUDA events) of "
                                        "MEASURED_PEAKS.json (%s); sustained figure %.1f -> frac %.3f" % (
                                            how, sustained, achieved / sustained)},
            "whole_path_tflops": value / world * FLOPS_PER_EVAL.get((args.blocks, args.hidden), 0) / 1e12,
            "wall_s_timed_region": wall,
        }
        if sp:
            line["selfplay"] = {"sims_per_s": float(tot[0].item()) / sp_seconds,
                                "nn_evals_per_s": float(tot[1].item()) / sp_seconds,
                                "games_per_gpu": args.selfplay_games, "sims_per_move": 200, "mode": "vanilla (config 5 rules: "
                                "Tier 1 off, material off), one leaf in flight per game; tree operations on the device "
                                "(cb_selfplay_run pipeline 3: select, forward, expand/play kernels in one CUDA graph)",
                                "seconds": sp_seconds, "mean_batch": sp["mean_batch"],
                                "forward_fraction": sp["eval_seconds"] / max(sp["seconds"], 1e-9),
                                "host_driven_sims_per_s": float(tot[2].item()) / max(sp_host_seconds, 1e-9),
                                "koth_800_sims_per_move_sims_per_s": float(tot[3].item()) / max(sp_koth_seconds, 1e-9),
                                "two_pools_sims_per_s": float(tot[4].item()) / max(sp_two_seconds, 1e-9),
                                "two_pools_material_pe_sims_per_s": float(tot[5].item()) / max(sp_two_pe_seconds, 1e-9),
                                "with_sample_records_sims_per_s": float(tot[6].item()) / max(sp_rec_seconds, 1e-9),
                                "sample_records_per_s": float(tot[7].item()) / max(sp_rec_seconds, 1e-9),
                                "two_pools_games_per_gpu": 2 * args.selfplay_games,
                                "two_pools_mode": "cb_selfplay_run pipeline 4: two pools of games_per_gpu games alternate, "
                                                  "the tree kernel of one beside the forward kernel of the other; "
                                                  "material_pe: dM by the principal-exchange line on the device",
                                "host_threads_per_gpu": nthreads}
        if gathered is not None:
            line["nccl_sample_gather_records"] = gathered
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = time_cpu_path(args, boards, q, idx, off, blob, 20.0, 3, 1)
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
