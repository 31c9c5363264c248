#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the feasibility-and-bound hot path.

Workload (BASELINE.json configs[1]): rect 16x16 ceiling, 1x1 supports, find the minimum support count.
One "step" = one epoch of the hot path on one GPU: the SLS kernel (b) advances every chain of the portfolio by
`--epoch-steps` steps, the best-reduce kernel folds the chains' best counts into the device-resident bound and (N > 1)
one NCCL all-reduce-min over NVLink shares that bound between the ranks.  value = candidate layouts evaluated per
second over all ranks (every candidate is scored exactly: its uncovered-tile count is computed, incrementally).

Beside it, on rank 0 at N = 1: time-to-optimal (fresh portfolio -> first layout with the proven optimum of 15),
the full-evaluation kernel (a) streaming 4 Mi candidate layouts from HBM, the CNF kernel (c), the measured
integer-issue / shared-memory peaks, the end-to-end number through the C ABI with host buffers, and a bounded CPU
baseline (the oracle's `validate`, i.e. the reference's own layout check, on the host cores).

`--impl reference` times the reference's CPU path instead (oracle port: `validate` throughput on all host threads
for the same metric, plus the CDCL bound-tightening loop's time to the same optimum) and never touches the GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate layouts evaluated/sec"
UNIT = "layouts/s"
WORKLOAD = "rect 16x16 ceiling, 1x1 supports, find minimum support count (BASELINE.json configs[1])"
OPTIMUM_RECT16 = 15   # UNSAT proven at <= 14 (re-derived by oracle CDCL: tests/test_oracle.py proves ex1-3; rect16 in DESIGN.md)
# algorithmic integer work (DESIGN.md "kernel (b)"): thread-ops, counted from the kernel's own counters
A_SCORE = 7 * 4 + 2      # per candidate scored: 7 window rows x (shift, and, pack, add) + key build
A_FLIP = 7 * 20          # per support added/removed: 7 window rows x (row mask 6 + five-plane add/sub 10 + derive 4)
KERNEL_NAMES = {1: "sls_kernel (one chain per warp)", 2: "sls_h16_kernel (two chains per warp)", 3: "sls_t16_kernel (one chain per thread)"}
# per-launch DRAM traffic and issue statistics of each variant from its committed ncu capture (profiles/)
KERNEL_NCU = {
    2: {"traffic": 3067392, "traffic_note": "dram bytes of one launch (ncu, 9472 chains x 512 steps): chain states only",
        "ncu": "ALU pipe 72% busy, 276 warp instructions per chain step (profiles/r1_sls_h16_kernel.md)"},
    3: {"traffic": 11065344, "traffic_note": "dram bytes of one launch (ncu, 56832 chains x 512 steps): chain states + site lists",
        "ncu": "ALU pipe 75.5% busy, issue slots 72% busy, 73 warp instructions per chain step, 17.8 shared-memory wavefronts per chain step (profiles/r1_sls_t16_kernel.md)"},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_validate_rate(grid, seconds=10.0, threads=None):
    """The reference's own layout check (PlatformLayout::validate, oracle port) on the host cores, bounded sample."""
    import oracle.oracle as O
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(0)
    n = 16384 * threads
    sites = (rng.random((n, grid.shape[0], grid.shape[1])) < 0.06).astype(np.uint8)
    done, t_total = 0, 0.0
    while t_total < seconds:
        _, _, sec = O.validate_sites_batch(grid, sites, threads=threads, flat=True)
        done += n
        t_total += sec
    return done / t_total, threads, (f"{done} random 1x1 layouts (6% density) on rect 16x16 through the oracle's flat-array port of "
                                     f"PlatformLayout::validate, {threads} threads, {t_total:.1f} s")


def cdcl_time_to_optimum(grid, optimum, conflict_budget=400000):
    """Oracle CDCL bound-tightening loop (Glucose stand-in): seconds until the first layout with `optimum` platforms."""
    import oracle.oracle as O
    r = O.solver_loop(grid, O.PLATFORMS_1X1, conflict_budget=conflict_budget)
    t = 0.0
    for s in r["steps"]:
        t += s["seconds"]
        if s["result"] == 10 and s["count"] <= optimum:
            return t, True
    return t, False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid = np.ones((16, 16), np.uint8)
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    rates = []
    sample = ""
    for i in range(args.warmup + args.steps):
        rate, threads, sample = cpu_validate_rate(grid, seconds=per_step)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    t_opt, reached = cdcl_time_to_optimum(grid, OPTIMUM_RECT16)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards / bool",
        "data": "synthetic", "config": {"workload": WORKLOAD, "note": "reference = CPU path (Rust + Glucose cannot be built here: oracle port)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "time_to_optimal_ms": t_opt * 1e3 if reached else None,
        "time_to_optimal_note": "oracle CDCL bound-tightening loop (Glucose stand-in), 1 thread, time to the first layout with 15 supports; UNSAT proof not included",
    }
    print(json.dumps(line), flush=True)


def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch, dist, world, rank, local


def comm_init(eng, torch, dist, rank, world):
    """The engine's own NCCL communicator: torch.distributed only carries the 128-byte id."""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)


def run_c5(args):
    """BASELINE.json configs[4]: batch of 100k synthetic 32x32 terrains (p = 0.7), contiguous shards per GPU, no data-path
    collective; one all-gather of the per-terrain counts at the end.  A step = one pass over the whole batch."""
    import ctypes as C
    torch, dist, world, rank, local = setup_dist()
    import timberborn_support_solver_b200 as T
    eng = T.Engine(local)
    n_total = args.terrains
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    lib = T.load()
    grids = np.zeros((hi - lo, 32, 32), np.uint8)
    for i, t in enumerate(range(lo, hi)):
        lib.tss_world_synthetic(32, 32, 1, t, int(0.7 * (1 << 24)), grids[i].ctypes.data_as(C.POINTER(C.c_uint8)))
    W, K = max(args.warmup, 1), args.steps
    for _ in range(W):
        eng.solve_batch(grids[: min(len(grids), 8192)], seed=1, steps=args.batch_steps)
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s0 = eng.stats()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(K):
        counts, layouts = eng.solve_batch(grids, seed=1, steps=args.batch_steps, want_layouts=True)
        dev_ms += eng.stats()["device_ms"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    s1 = eng.stats()
    sampler.stop_flag = True
    sampler.join()
    # final gather: 4 bytes per terrain
    cd = torch.from_numpy(counts).cuda()
    if world > 1:
        sizes = [(r + 1) * n_total // world - r * n_total // world for r in range(world)]
        parts = [torch.empty(sz, dtype=torch.int32, device="cuda") for sz in sizes]
        dist.all_gather(parts, cd)
        allc = torch.cat(parts)
    else:
        allc = cd
    tt = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(s1["candidates_scored"] - s0["candidates_scored"]), float(s1["kernel_launches"] - s0["kernel_launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        dev_ms_max, wall_ms_max = (float(x) for x in tt.tolist())
        tiles = grids.reshape(len(grids), -1).sum(1)
        # every reported layout of rank 0's shard re-evaluated by kernel (a) in per-terrain mode: complete, count matches
        terr_rows = (grids.astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(2, dtype=np.uint32)
        g_dev, l_dev = torch.from_numpy(terr_rows.view(np.int32)).cuda(), torch.from_numpy(layouts.view(np.int32)).cuda()
        out = torch.empty((len(grids), 2), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        eng.eval_compact_dev(g_dev.data_ptr(), 32, 32, l_dev.data_ptr(), len(grids), out.data_ptr(), per_layout_terrain=True)
        torch.cuda.synchronize()
        res = out.cpu().numpy()
        ok = bool((res[:, 0] == 0).all() and np.array_equal(res[:, 1], counts))
        line = {"metric": "terrains solved/sec (per-terrain best support count)", "value": n_total * K / (dev_ms_max * 1e-3), "unit": "terrains/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32 bitboards / bool", "data": "synthetic",
                "config": {"workload": f"batch of {n_total} synthetic 32x32 terrains (p=0.7), 1x1 supports (BASELINE.json configs[4])", "sls_steps_per_chain": args.batch_steps,
                           "chains_per_terrain": 4, "parallelism": f"terrain shards x{world}, no data-path collective, final all-gather of counts",
                           "l2": "inputs larger than L2 (100 MB of terrains, 800 MB of reach tables per pass)"},
                "e2e": {"value": n_total * K / (wall_ms_max * 1e-3), "unit": "terrains/s", "h2d_bytes_per_step": int(grids.size), "d2h_bytes_per_step": int(counts.nbytes + layouts.nbytes),
                        "note": "tss_solve_batch from host u8 grids to host counts + layouts, wall clock, slowest rank"},
                "gpu_launches": int(tot[1].item()), "candidates_per_s": float(tot[0].item()) / (dev_ms_max * 1e-3), "mean_count": float(allc.float().mean().item()),
                "mean_ceiling_tiles": float(tiles.mean()), "witnesses_revalidated_by_kernel_a": ok, "clocks": sampler.summary(),
                }
        if world == 1:
            # bounded CPU baseline: the reference's loop (oracle CDCL as the Glucose stand-in, 1 thread per instance like the
            # reference) on the first terrains of the batch, a conflict budget per solve instead of minutes per terrain
            import oracle.oracle as O
            t0 = time.perf_counter()
            cpu_counts = []
            for t in range(2):
                r = O.solver_loop(grids[t], O.PLATFORMS_1X1, conflict_budget=50000)
                cpu_counts.append(min(s["count"] for s in r["steps"] if s["result"] == 10))
            cpu_s = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 2 / cpu_s, "unit": "terrains/s", "cores": 1, "kind": "port",
                                    "sample": f"terrains 0-1 of the batch through the oracle's CDCL bound-tightening loop, 50000 conflicts per solve, {cpu_s:.1f} s: "
                                              f"counts {cpu_counts} (unproven), GPU counts for the same terrains {[int(c) for c in counts[:2]]}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_c4(args):
    """BASELINE.json configs[3]: synthetic 256x256 ceiling (p = 0.7), window-decomposed SLS portfolio, one seed set per GPU,
    all-reduce-min of (count, rank) + the winner's layout after every phase.  A step = one phase."""
    torch, dist, world, rank, local = setup_dist()
    import timberborn_support_solver_b200 as T
    eng = T.Engine(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        comm_init(eng, torch, dist, rank, world)
    g = T.WorldGrid.synthetic(256, 256, 1, 0)
    s = eng.search(g, seed=1, n_chains=args.chains, chain_offset=rank * 1000000)     # 0 = one wave of chains over the windows
    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        s.run(args.phase_steps, 0)
    s.best_count()
    s0 = eng.stats()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        s.run(args.phase_steps, 0)
        b.record()
    torch.cuda.synchronize()
    count = s.global_best()
    s1 = eng.stats()
    sampler.stop_flag = True
    sampler.join()
    lay = s.best_layout()            # re-validated by kernel (a) inside the engine
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # e2e: the same terrain through the one-shot C-ABI call with HOST buffers on every rank (workspace creation, terrain
    # upload, K phases, layout download + validation inside the timed region)
    e0 = eng.stats()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res, lay_e2e = eng.solve_upper_bound(g, card_limit=None, seed=50 + rank, max_steps=K * args.phase_steps)
    t_e2e = time.perf_counter() - t0
    e1 = eng.stats()
    e2e_t = torch.tensor([float(e1["candidates_scored"] - e0["candidates_scored"]), t_e2e], dtype=torch.float64, device="cuda")
    e2e_max = e2e_t.clone()
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
        dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(s1["candidates_scored"] - s0["candidates_scored"]), float(s1["kernel_launches"] - s0["kernel_launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(tt.item())
        line = {"metric": METRIC, "value": float(tot[0].item()) / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards / bool", "data": "synthetic",
                "config": {"workload": "synthetic 256x256 random ceiling mask (density 0.7), 1x1 supports, SLS portfolio with all-reduce-min bound (BASELINE.json configs[3])",
                           "phase_steps": args.phase_steps, "chains_total": s.n_chains, "parallelism": f"portfolio x{world}: window decomposition per GPU, winner's layout shipped after every phase"},
                "gpu_launches": int(tot[1].item()), "best_count": count, "layout_platforms": lay.platform_count(), "ceiling_tiles": int(g.data.sum()),
                "phases_total": W + K, "clocks": sampler.summary(),
                "e2e": {"value": float(e2e_t[0].item()) / float(e2e_max[1].item()), "unit": UNIT, "h2d_bytes_per_step": int(g.data.size // max(K, 1)),
                        "d2h_bytes_per_step": int(20 * lay_e2e.platform_count() // max(K, 1)), "count": lay_e2e.platform_count(),
                        "note": f"tss_solve_upper_bound from a host u8 grid to a host platform list on each of the {world} rank(s): workspace creation, upload, "
                                f"{K} phases of {args.phase_steps} steps, layout download and validation; wall clock, sum over ranks / slowest rank"}}
        print(json.dumps(line), flush=True)
    s.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--epoch-steps", type=int, default=4096, help="SLS steps per chain per bench step")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0 = fill the device: 3 CTAs x 128 one-thread chains per SM)")
    ap.add_argument("--kernel", type=int, default=0, help="SLS kernel variant (tss.h TSS_KERNEL_*: 0 auto, 1 warp, 2 half-warp, 3 thread)")
    ap.add_argument("--quick", action="store_true", help="skip the side measurements (peaks, eval/cnf kernels, cpu baseline)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"], help="c2 = the bench line (rect 16x16); c4 / c5 = the other named configs, for context")
    ap.add_argument("--terrains", type=int, default=100000, help="c5: terrains in the batch")
    ap.add_argument("--batch-steps", type=int, default=2000, help="c5: SLS steps per chain")
    ap.add_argument("--phase-steps", type=int, default=4000, help="c4: SLS steps per window-decomposition phase")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_c5(args)
    if args.workload == "c4":
        return run_c4(args)

    import torch
    import torch.distributed as dist

    import timberborn_support_solver_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps

    eng = T.Engine(local)
    # torch events only see torch's current stream: make one explicit (non-default) stream current and hand it to the
    # engine, so every kernel of the hot path and every event of this file live on the same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    info = eng.device_info()
    grid = T.WorldGrid(np.ones((16, 16), np.uint8))
    exchange = "none (single GPU)"
    if world > 1:
        # the path's one real exchange: an all-reduce-min of the best-known count (4 bytes, latency bound).  The engine does
        # it itself, in-stream on the device-resident bound (ncclAllReduce inside tss_search_run); torch.distributed only
        # carries the 128-byte NCCL id from rank 0 to the other ranks.
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        exchange = "ncclAllReduce(min, 1 x int32) in-stream inside tss_search_run, bound stays in HBM"
    n_chains = args.chains
    if not n_chains:                 # engine default: fills the device for the chosen kernel
        probe = eng.search(grid, kernel=args.kernel)
        n_chains = probe.n_chains
        probe.close()
    search = eng.search(grid, seed=1, n_chains=n_chains, chain_offset=rank * n_chains, kernel=args.kernel)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step():
        search.run(args.epoch_steps, 0)

    # mean |R(s)| of the terrain (SURVEY.md §8d asks for it beside the rate): one-support layouts through kernel (a)
    one_hot = np.eye(grid.data.size, dtype=np.uint8).reshape(-1, grid.height, grid.width)
    unc1, _ = eng.eval_sites(grid, one_hot)
    mean_reach = float((int(grid.data.sum()) - unc1[grid.data.reshape(-1) != 0]).mean())
    for _ in range(W):
        step()
    search.best_count()        # synchronises and folds the warm-up's device counters into the stats BEFORE the baseline snapshot
    torch.cuda.synchronize()
    s0 = eng.stats()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    sampler.join()
    best = search.best_count()
    s1 = eng.stats()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    scored = s1["candidates_scored"] - s0["candidates_scored"]
    sls_steps = s1["sls_steps"] - s0["sls_steps"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    tot = torch.tensor([float(scored), float(sls_steps), float(launches)], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    scored_all, steps_all, launches_all = (float(x) for x in tot.tolist())
    ms_total = float(tmax.item())
    value = scored_all / (ms_total * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards / bool", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_per_gpu": n_chains, "epoch_steps": args.epoch_steps, "parallelism": f"portfolio x{world} (independent seeds, all-reduce-min of the bound per step)", "exchange": exchange,
                   "l2": "flushed between timed steps (256 MiB memset outside the event pairs); the kernel's working set is registers + the CTA's shared memory (boards, reach table), HBM is touched at epoch start/end only"},
        "gpu_launches": int(launches_all), "best_count": best, "sls_steps_per_s": steps_all / (ms_total * 1e-3),
        "flips_per_s": 2.0 * steps_all / (ms_total * 1e-3), "candidates_per_step": scored_all / max(steps_all, 1.0), "mean_reach": mean_reach,
        "wall_ms_total": t_wall * 1e3, "clocks": sampler.summary(),
    }

    if rank == 0:
        # ---------------- roofline of the dominant kernel (SLS): algorithmic integer thread-ops / event time vs measured LOP3 issue peak
        pk = eng.measure_peaks()
        flips = 2.0 * steps_all                      # a swap step removes one support and adds one
        int_ops = A_SCORE * scored_all + A_FLIP * flips
        achieved = int_ops / (ms_total * 1e-3) / 1e9 / world
        variant = args.kernel or (3 if n_chains >= info["sm_count"] * 48 else 2)      # engine's auto rule for a 16x16 grid (engine.cu search_kernel)
        ncu = KERNEL_NCU.get(variant, {"traffic": None, "traffic_note": "no capture for this variant", "ncu": ""})
        line["roofline"] = {"bound": "int_issue", "achieved": achieved, "peak": pk["lop3_gops"], "unit": "Gop/s", "frac": achieved / pk["lop3_gops"],
                            "traffic": ncu["traffic"], "traffic_note": ncu["traffic_note"], "kernel": KERNEL_NAMES[variant], "ncu": ncu["ncu"],
                            "algorithmic_ops": f"{A_SCORE} thread-ops per candidate scored + {A_FLIP} per support added/removed (DESIGN.md kernel (b))",
                            "peak_source": "measured in this run (tss_measure_peaks: dependent-free LOP3 chains at full occupancy)",
                            "note": "no dense contraction and ~0 HBM traffic in the step loop: the bound is integer issue (SURVEY.md §8d); per-GPU figures"}
        line["measured_peaks"] = pk
    # ---------------- e2e through the C ABI with HOST buffers (grid in, layout out) on every rank at once: a fixed step budget
    # per call; terrain upload, reach table, epochs with their host round trips, witness validation and the layout copy back
    # are all inside the timed region.  Whole-job value = candidates of all ranks / slowest rank's wall time.
    steps_per_call, n_calls = 16384, 5
    eng.solve_upper_bound(grid, card_limit=None, seed=199, max_steps=steps_per_call)      # workspace warm-up (allocation)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = eng.stats()
    t0 = time.perf_counter()
    for i in range(n_calls):
        res, lay = eng.solve_upper_bound(grid, card_limit=None, seed=200 + 16 * rank + i, max_steps=steps_per_call)
    t_e2e = time.perf_counter() - t0
    e1 = eng.stats()
    e2e_t = torch.tensor([float(e1["candidates_scored"] - e0["candidates_scored"]), t_e2e], dtype=torch.float64, device="cuda")
    e2e_max = e2e_t.clone()
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
        dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    line["e2e"] = {"value": float(e2e_t[0].item()) / float(e2e_max[1].item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(grid.data.size + 8), "d2h_bytes_per_step": int(20 * lay.platform_count() + 16 * 9 + 320),
                   "note": f"tss_solve_upper_bound from a host u8 grid to a host platform list on each of the {world} rank(s), {steps_per_call} steps/chain per call, "
                           f"{n_calls} calls, wall clock incl. copies, epoch round trips and witness validation; sum over ranks / slowest rank"}
    if rank == 0:
        # ---------------- time-to-optimal: fresh portfolio -> first layout with 15 supports (incl. host round trips); one GPU finds
        # it in a fraction of a millisecond, so this is per rank and identical at every N (the one-shot solve never communicates)
        tto = []
        for seed in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res, lay = eng.solve_upper_bound(grid, card_limit=OPTIMUM_RECT16, seed=100 + seed)
            tto.append((time.perf_counter() - t0) * 1e3)
            assert res == T.SAT and lay.platform_count() == OPTIMUM_RECT16
        line["time_to_optimal_ms"] = float(np.median(tto[2:]))
        line["time_to_optimal_note"] = ("tss_solve_upper_bound(card_limit=15) from host buffers on one GPU: terrain upload, reach table, epochs of 64.. "
                                        "steps, witness re-validated by kernel (a); median of 5 calls after 2 warm-up calls")
    if rank == 0 and not args.quick and world == 1:
        hbm_peak, hbm_src = peaks()
        # ---------------- kernel (a): stream 4 Mi candidate layouts (32 B each, 128 MiB > L2) from HBM
        n_lay = 4 << 20
        lay_dev = torch.randint(0, 1 << 16, (n_lay, 16), dtype=torch.int32, device="cuda")
        lay_dev = (lay_dev & torch.randint(0, 1 << 16, (n_lay, 16), dtype=torch.int32, device="cuda") & torch.randint(0, 1 << 16, (n_lay, 16), dtype=torch.int32, device="cuda")
                   & torch.randint(0, 1 << 16, (n_lay, 16), dtype=torch.int32, device="cuda")).to(torch.int16).contiguous()   # ~6% density
        grid_dev = torch.full((16,), -1, dtype=torch.int16, device="cuda")
        out_dev = torch.empty((n_lay, 2), dtype=torch.int32, device="cuda")
        for _ in range(3):
            eng.eval_compact_dev(grid_dev.data_ptr(), 16, 16, lay_dev.data_ptr(), n_lay, out_dev.data_ptr())
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.eval_compact_dev(grid_dev.data_ptr(), 16, 16, lay_dev.data_ptr(), n_lay, out_dev.data_ptr())
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        gbs = n_lay * (32 + 8) / (ms * 1e-3) / 1e9
        line["eval_kernel"] = {"layouts_per_s": n_lay / (ms * 1e-3), "ms": ms, "n": n_lay, "bytes_per_layout": 40,
                               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src},
                               "int_gops": 19 * 16 * n_lay / (ms * 1e-3) / 1e9, "int_note": "A_eval = 19 ops x 16 row words per layout (SURVEY.md §8d accounting)"}
        # ---------------- kernel (c): CNF check of 8192 witness assignments against the encoder's clauses (incl. the totalizer
        # of the at-most-15 bound): the SLS witness, completed by unit propagation, replicated; every 64th copy has one
        # support removed (those must come back falsified)
        enc = T.Encoding.encode(T.PLATFORMS_DEFAULT[:1], grid)
        cnf = enc.with_limits(T.PlatformLimits.new_unweighted({T.PlatformDef(1, 1): 15}))
        dev = eng.upload_cnf(cnf)
        res, wit = eng.solve_upper_bound(grid, card_limit=OPTIMUM_RECT16, seed=300)
        full = np.full((1, cnf.n_vars + 1), 2, np.uint8)
        base_a = eng.layout_to_assignment(enc, wit)
        full[0, : len(base_a)] = base_a
        prop, conflict, rounds = dev.propagate(full)
        prop[prop == 2] = 0
        assert conflict[0] < 0 and dev.check(prop)[0][0] == 0
        a = np.repeat(prop, 131072, axis=0)     # 4096 words per variable and polarity: long enough a launch (0.17 ms) to time the kernel, not its launch
        first_support = int(enc.vars().plat_var[[p.y * 16 + p.x for p in wit.platforms().values()][0], 0])
        a[::64, first_support] = 0
        nf, _ = dev.check(a)
        nf, _ = dev.check(a)
        assert (nf[::64] > 0).all() and nf.reshape(-1, 64)[:, 1:].sum() == 0
        cnf_ms = eng.stats()["device_ms"]
        nbw = (len(a) + 31) // 32
        cnf_bytes = 2 * (cnf.n_vars + 1) * nbw * 4 + 8 * len(a)      # both bit-sliced planes read once + (count, first) per assignment written
        cnf_ops = len(cnf.lits) * nbw                                # one logic op per literal per 32 assignments (DESIGN.md A_cnf): thread-ops
        cnf_reads = len(cnf.lits) * nbw * 4                          # ... and one 4-byte plane word per literal and word (L2 traffic: a variable occurs in ~3 clauses)
        line["cnf_kernel"] = {"clause_evals_per_s": cnf.n_clauses * len(a) / (cnf_ms * 1e-3), "clauses": cnf.n_clauses, "ms": cnf_ms,
                              "roofline": {"bound": "hbm", "achieved": cnf_bytes / (cnf_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": cnf_bytes / (cnf_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                                           "plane_reads_gbs": cnf_reads / (cnf_ms * 1e-3) / 1e9, "int_gops": cnf_ops / (cnf_ms * 1e-3) / 1e9,
                                           "note": "algorithmic bytes = both bit-sliced assignment planes once + results; the planes (84 MB) are L2 resident at this size "
                                                   "and every literal reads one plane word, so the kernel runs on L2 bandwidth and load latency (plane_reads_gbs), with one logic op per word"},
                              "literals": int(len(cnf.lits)), "assignments": len(a), "propagation_rounds": rounds,
                              "input": "SLS witness completed by unit propagation x 131072, every 64th with one support removed"}
        # ---------------- the other named configs, for context (parity-test cases, not the bench line): wall clock through the C ABI
        others = {}
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures.json")))
        ex1_rows = fx["ex1"]["grid"]   # test/ex1.toml
        ex1 = T.WorldGrid.from_toml("[world]\ngrid = [\n" + "".join(f'    "{r}",\n' for r in ex1_rows) + "]\n")
        t1s = []
        for i in range(7):
            t0 = time.perf_counter()
            res, lay3 = eng.solve_upper_bound(ex1, T.PLATFORMS_DEFAULT, card_limit=1, seed=1 + i)
            t1s.append((time.perf_counter() - t0) * 1e3)
            assert res == T.SAT and lay3.platform_count() == 1
        others["C1 ex1 default-8 (REPL set)"] = {"count": lay3.platform_count(), "proven_optimum": 1, "ms": float(np.median(t1s[2:])),
                                                 "note": "tss_solve_upper_bound(card_limit=1) from host buffers, median of 5 calls after 2 warm-up calls"}
        ex2 = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(max(len(q) for q in fx["ex2"]["grid"]))] for r in fx["ex2"]["grid"]], np.uint8)
        ex2[8:10, 8:10] = 1          # README terrain = test/ex2.toml with (8,8),(9,8),(8,9),(9,9) set to ceiling
        t3s = []
        for i in range(7):
            t0 = time.perf_counter()
            res, lay2 = eng.solve_upper_bound(T.WorldGrid(ex2), card_limit=14, seed=40 + i)
            t3s.append((time.perf_counter() - t0) * 1e3)
            assert res == T.SAT and lay2.platform_count() == 14
        others["C3 README 21x16 terrain, 1x1 supports"] = {"count": 14, "proven_optimum": 14, "ms": float(np.median(t3s[2:])),
                                                         "note": "tss_solve_upper_bound(card_limit=14) from host buffers, median of 5 calls after 2 warm-up calls; the README transcript "
                                                                 "stops at 15 (README.md:117-119); the oracle's CDCL loop (Glucose stand-in, 1 thread, build container) finds 14 after 1.2 s and proves it 10 s later"}
        g4 = T.WorldGrid.synthetic(256, 256, 1, 0)
        s4 = eng.search(g4, seed=1)          # default: one wave of chains over the windows
        t0 = time.perf_counter()
        for _ in range(24):
            s4.run(4000, 0)
        c4 = s4.best_count()
        others["C4 256x256 p=0.7 (window decomposition, 24 phases x 4000 steps)"] = {"count": c4, "ceiling_tiles": int(g4.data.sum()), "ms": (time.perf_counter() - t0) * 1e3}
        s4.close()
        n5 = 16384
        g5 = np.stack([T.WorldGrid.synthetic(32, 32, 1, t).data for t in range(n5)])
        t0 = time.perf_counter()
        c5 = eng.solve_batch(g5, seed=1, steps=2000)
        dt5 = time.perf_counter() - t0
        others["C5 batch of 32x32 p=0.7 terrains (16384 of the 100k, 2000 steps x 4 chains each)"] = {
            "terrains_per_s": n5 / dt5, "mean_count": float(c5.mean()), "ms": dt5 * 1e3,
            "note": "host Glucose stand-in needs minutes per terrain (3 sampled terrains: best 76/78/73 after 5 min each, GPU 73/73/67)"}
        line["other_configs"] = others
        # ---------------- CPU baseline (bounded sample, rank 0, N = 1)
        rate, threads, sample = cpu_validate_rate(grid.data, seconds=10.0)
        t_opt, reached = cdcl_time_to_optimum(grid.data, OPTIMUM_RECT16)      # the bound-tightening loop itself, 1 thread like the reference
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                "time_to_optimal_ms": t_opt * 1e3 if reached else None,
                                "time_to_optimal_note": "oracle CDCL bound-tightening loop (Glucose stand-in, crates/repl/src/main.rs:280-366), 1 solver thread as in the reference "
                                                        "(solver_runner.rs:15), time to the first layout with 15 supports on rect 16x16; the UNSAT proof of 14 (minutes) is not included"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    search.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
