#!/usr/bin/env python
"""bench.py — simulated IRV elections / second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3] [--impl ours|reference]

A *step* is one pass of the hot path (Dirichlet-tree posterior draw -> IRV -> win counts) over one
batch of simulated elections of the named configuration, `--elections-per-gpu` per GPU (weak
scaling; the default C3 step is BASELINE's 100k elections on every GPU).
Ranks simulate disjoint global election ranges, then all-reduce the <=32 win counts (NCCL).

 value      tree already resident in HBM, win counts left on the device; CUDA events on the launch
            stream, max over ranks.
 e2e        the same batch through the host-buffer C-ABI call (dtree_sample_posterior) with the
            device mirror invalidated first: CSR flatten + H2D of the tree, kernel, D2H of the counts.
 roofline   simulate_kernel against measured HBM bandwidth, with SURVEY §8(d)'s algorithmic bytes
            per election taken from the CPU oracle's counters on the same inputs.
 cpu_baseline / --impl reference
            the reference's own C++ (oracle/_ref, else the port) on all host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

METRIC = "simulated IRV elections/sec"
SEED = "BENCHSEEDx"
DEFAULT_EPG = {"C1": 4_000_000, "C2": 1_000_000, "C3": 100_000, "C4": 4_000, "C5": 200_000}


def describe(name, cfg, epg):
    return ("%s: %d candidates, %d observed ballots (%s), n_ballots=%d, min_depth=%d, max_depth=%d, a0=%g, vd=%s; "
            "%d elections per GPU per step" % (name, cfg["n_candidates"], cfg["observed"][2 if cfg["observed"][0] == "pl" else 1],
                                               "Plackett-Luce gamma=%g seed %d" % (cfg["observed"][1], cfg["observed"][3])
                                               if cfg["observed"][0] == "pl" else "Wakehurst sample seed %d" % cfg["observed"][2],
                                               cfg["n_ballots"], cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], epg))


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_oracle():
    from oracle import oracle
    kind = "reference" if oracle.available("reference") else "port"
    return oracle, oracle.load(kind), kind


def cpu_run(olib, cfg, prefs, offsets, n_elections, threads, seed):
    t = olib.tree(cfg["n_candidates"], cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"])
    t.update((prefs, offsets))
    freq, secs = t.sample_posterior(n_elections, cfg["n_ballots"], 1, False, threads, seed)
    t.close()
    return freq, secs


def cpu_sample_size(cfg, threads):
    """Elections for a bounded CPU sample: ~10-30 s of work on `threads` cores (per-election cost from BASELINE.md)."""
    per_election = {4: 3e-5, 6: 1.2e-3, 8: 0.06, 10: 0.9, 20: 10.0}.get(cfg["n_candidates"], 0.5)
    n = int(15.0 * threads / per_election)
    return max(2 * threads, min(n, cfg["n_elections"]))


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args, cfg, name):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg_wl = load_pkg().workloads
    oracle, olib, kind = load_oracle()
    threads = host_threads()
    prefs, offsets = pkg_wl.observed_ballots(name)
    per_step = max(4 * threads, cpu_sample_size(cfg, threads) // 2)
    t = olib.tree(cfg["n_candidates"], cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"])
    t.update((prefs, offsets))
    secs = []
    for i in range(args.warmup + args.steps):
        _, s = t.sample_posterior(per_step, cfg["n_ballots"], 1, False, threads, SEED + str(i))
        if i >= args.warmup:
            secs.append(s)
    total = float(sum(secs))
    value = per_step * args.steps / total
    sample = "%d elections per step (of %d), %d threads, thread pool of R_tree.cpp:213-242" % (per_step, cfg["n_elections"], threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "elections/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(name, cfg, per_step), "note": "CPU only; n_gpus is the launch shape, rank 0 does all the work"},
        "cpu_baseline": {"value": value, "unit": "elections/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "elections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--elections-per-gpu", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--urn-max", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-mode", action="store_true", help="do not also time the other kernel")
    ap.add_argument("--mode", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="1: deferred expansion kernel, 0: materialising kernel, -1: the library's choice")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    pkg = load_pkg()
    W = pkg.workloads
    name = args.config
    cfg = W.CONFIGS[name]
    if args.impl == "reference":
        return run_reference(args, cfg, name)

    import torch
    import torch.distributed as dist
    from elections_dtree_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    epg = args.elections_per_gpu or DEFAULT_EPG[name]
    nc = cfg["n_candidates"]
    prefs, offsets = W.observed_ballots(name)
    t = pkg.dirichlet_tree(W.candidate_names(nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=local)
    t.update_indices(prefs, offsets)
    if args.block_threads or args.blocks_per_sm:
        t.tune(block_threads=args.block_threads, blocks_per_sm=args.blocks_per_sm)
    if args.urn_max >= 0:
        t.tune(urn_max=args.urn_max)
    t.tune(mode=args.mode)
    L = N.lib()
    N.check(L.dtree_sync_device(t._h))
    seed = SEED.encode()
    stream = torch.cuda.current_stream(dev)
    wins = torch.zeros(nc, dtype=torch.int64, device=dev)

    total = torch.zeros(nc, dtype=torch.int64, device=dev)
    # Larger than the 126 MB L2: zeroed between timed steps so that no step starts with the previous one's lines.
    l2_flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(mode, n_el, steps, warmup, offset):
        """K timed steps of n_el elections per GPU with the tree resident and the counts left on the device."""
        t.tune(mode=mode)

        def launch(i):
            # global election ids: step i covers [i*n_el*world, (i+1)*n_el*world); this rank's contiguous share
            first = offset + (i * world + rank) * n_el
            wins.zero_()
            N.check(L.dtree_sample_posterior_device(t._h, first, n_el, cfg["n_ballots"], 1, 0, seed, len(seed),
                                                    wins.data_ptr(), stream.cuda_stream))

        for i in range(warmup):
            launch(i)
            if world > 1:
                dist.all_reduce(wins, op=dist.ReduceOp.SUM)
        N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        # finish() lets the library resize its per-election scratch from what the warm-up elections used; one more
        # untimed launch makes that (re)allocation happen before the timed region rather than inside its first step
        for _ in range(2):
            launch(warmup - 1)
            if world > 1:
                dist.all_reduce(wins, op=dist.ReduceOp.SUM)
            N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ev = []
        total.add_(wins)  # first use of these torch ops loads their kernels: not inside the timed region
        l2_flush.zero_()
        total.zero_()
        torch.cuda.synchronize(dev)
        ev0.record(stream)
        step_ev = []
        for i in range(steps):
            l2_flush.zero_()  # between timed steps, not inside one
            s0, k0, k1, s1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            s0.record(stream)
            k0.record(stream)
            launch(warmup + i)
            k1.record(stream)
            kernel_ev.append((k0, k1))
            if world > 1:
                dist.all_reduce(wins, op=dist.ReduceOp.SUM)
            total.add_(wins)
            s1.record(stream)
            step_ev.append((s0, s1))
        ev1.record(stream)
        barrier()
        N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        st = t.last_stats()
        # a step = kernel(s) + all-reduce + accumulation; the K steps are timed one by one because the L2 flush sits
        # between them; the job's time is the sum, max over ranks
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in step_ev), ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms, ms_with_flush = float(ms[0].item()), float(ms[1].item())
        assert int(total.sum().item()) == n_el * world * steps
        return {"value": n_el * world * steps / (ms / 1e3), "elapsed_ms": ms, "elapsed_ms_with_flush": ms_with_flush,
                "steps": steps, "elections_per_gpu": n_el,
                "kernel_ms": float(np.mean([a.elapsed_time(b) for a, b in kernel_ev])),
                "kernel_ms_steps": [round(a.elapsed_time(b), 3) for a, b in kernel_ev], "stats": st,
                "step_ms_steps": [round(a.elapsed_time(b), 3) for a, b in step_ev],
                "after_kernel_ms_steps": [round(k[1].elapsed_time(s_[1]), 3) for k, s_ in zip(kernel_ev, step_ev)],
                "kernel": {0: "simulate_kernel", 1: "deferred_kernel", 2: "lane_kernel"}[st["mode"]]}

    sampler = ClockSampler(local) if rank == 0 else None
    main = measure(args.mode, epg, args.steps, args.warmup, 0)
    stats = main["stats"]
    elapsed_ms, kernel_ms, value = main["elapsed_ms"], main["kernel_ms"], main["value"]
    # the other kernel on the same inputs (fewer elections when it is the slower, materialising one)
    other = None
    if not args.single_mode:
        om = 0 if stats["mode"] >= 1 else 1
        o_epg = max(296, epg // 8) if om == 0 else epg
        other = measure(om, o_epg, max(2, min(args.steps, 3)), 3, 1 << 40)
    t.tune(mode=args.mode)

    # ---- e2e: host buffers through the C ABI, mirror invalidated so that the tree upload is inside -------------
    e2e_steps = max(1, args.steps)  # all K steps: host-side hiccups on a shared box are diluted, and visible in ms_steps
    for i in range(2):  # untimed warm-up of this path too (it lets the library settle its scratch sizing)
        N.check(L.dtree_invalidate_device(t._h))
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed=SEED, first_election=(1 << 41) + (i * world + rank) * epg)
    h2d0 = t.last_stats()["h2d_bytes"]
    barrier()
    t0 = time.perf_counter()
    e2e_ms = []
    for i in range(e2e_steps):
        ts = time.perf_counter()
        N.check(L.dtree_invalidate_device(t._h))
        first = ((args.warmup + args.steps + i) * world + rank) * epg
        w = t.sample_posterior_counts(epg, cfg["n_ballots"], seed=SEED, first_election=first)
        if world > 1:
            wt = torch.from_numpy(w.astype(np.int64)).to(dev)
            dist.all_reduce(wt, op=dist.ReduceOp.SUM)
            w = wt.cpu().numpy()
        e2e_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
    barrier()
    e2e_s = time.perf_counter() - t0
    e2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_s = float(e2.item())
    h2d = (t.last_stats()["h2d_bytes"] - h2d0) // e2e_steps
    e2e = {"value": epg * world * e2e_steps / e2e_s, "unit": "elections/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(nc * 8), "steps": e2e_steps, "ms_steps": e2e_ms}
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline denominators from the CPU oracle's counters on the same inputs -----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # the kernel's own per-election counters (exact for what was launched)
    # the materialising kernel's own counters when it ran (S, E, R describe materialisation), else the headline's
    mst = stats
    for cand in (main, other):
        if cand is not None and cand["kernel"] == "simulate_kernel":
            mst = cand["stats"]
    S_k = mst["slots"] / max(mst["elections"], 1)
    E_k = mst["entries"] / max(mst["elections"], 1) + (len(t.observed_indices()[2]))
    R_k = mst["irv_rereads"] / max(mst["elections"], 1)
    cpu_baseline = None
    b_alg = 4 * S_k + 32 * E_k + 16 * R_k + 8 * nc
    denominators = {"source": "simulate_kernel counters" if mst["mode"] == 0 else "deferred_kernel counters (not 8(d)'s)",
                    "S": S_k, "E": E_k, "R": R_k}
    if not args.no_cpu_baseline:
        oracle, olib, kind = load_oracle()
        port = oracle.load("port")
        ot = port.tree(nc, cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"])
        ot.update((prefs, offsets))
        n_stat = 3 if nc >= 10 else 20
        st = [ot.election_stats(cfg["n_ballots"], False, 1000 + i) for i in range(n_stat)]
        ot.close()
        S, E, R = (float(np.mean([s[k] for s in st])) for k in ("S", "E", "R"))
        b_alg = 4 * S + 32 * E + 16 * R + 8 * nc
        denominators = {"source": "oracle counters, mean of %d elections" % n_stat, "S": S, "E": E, "R": R,
                        "V_gamma": float(np.mean([s["V_gamma"] for s in st])),
                        "kernel_counters": {"S": S_k, "E": E_k, "R": R_k,
                                            "gammas": stats["gammas"] / max(stats["elections"], 1),
                                            "binomial_attempts": stats["binomial_attempts"] / max(stats["elections"], 1),
                                            "items": stats["items"] / max(stats["elections"], 1),
                                            "urn_items": stats["urn_items"] / max(stats["elections"], 1)}}
        if other is not None and other["kernel"] == "simulate_kernel":
            denominators["kernel_counters"]["items_full"] = other["stats"]["items"] / max(other["stats"]["elections"], 1)
        threads = host_threads()
        n_cpu = cpu_sample_size(cfg, threads)
        _, secs = cpu_run(olib, cfg, prefs, offsets, n_cpu, threads, SEED)
        cpu_baseline = {"value": n_cpu / secs, "unit": "elections/s", "cores": threads, "kind": kind,
                        "sample": "%d elections of the same workload in %.1f s, %d threads (R_tree.cpp:213-242 pool)" % (n_cpu, secs, threads)}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        traffic_tab = {}

    def roofline(m):
        """algorithmic bytes per launch / CUDA-event duration of the launch, against measured HBM copy
        bandwidth. The materialising kernel moves 8(d)'s bytes (B_alg = 4S + 32E + 16R + 8nc). The deferred kernel
        and lane kernels never create most of them, so they are rated on the bytes THEIR algorithm must move (4 B per outcome slot read +
        every 32 B node record written to or read back from the per-election arena, counted by the kernel, + 8nc);
        the 8(d) rate is kept beside it."""
        n_el, secs = m["elections_per_gpu"], m["kernel_ms"] / 1e3
        st = m["stats"]
        per_el = max(st["elections"], 1)
        if m["kernel"] == "simulate_kernel":
            bytes_el = b_alg
        else:
            # 4 B per outcome slot read + every 32 B node record the kernel wrote to / read back from its arena + 8nc
            bytes_el = 4.0 * st["slots"] / per_el + 32.0 * (st["entries"] + st["irv_rereads"]) / per_el + 8 * nc
        ach = bytes_el * n_el / secs / 1e9
        tr = traffic_tab.get(m["kernel"], {})
        out = {"kernel": m["kernel"], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
               "traffic": tr["bytes_per_election"] * n_el if tr.get("config") == name else None,
               "traffic_source": tr.get("capture") if tr.get("config") == name else None,
               "peak_source": peak_src, "algorithmic_bytes_per_election": bytes_el,
               "kernel_ms_per_launch": m["kernel_ms"], "kernel_ms_steps": m["kernel_ms_steps"],
               "step_ms_steps": m.get("step_ms_steps"), "after_kernel_ms_steps": m.get("after_kernel_ms_steps"),
               "elections_per_launch": n_el}
        if m["kernel"] != "simulate_kernel":
            eq = b_alg * n_el / secs / 1e9
            out["survey_8d"] = {"bytes_per_election": b_alg, "equivalent_GBps": eq, "equivalent_frac": eq / peak,
                                "note": "what a kernel that materialises every ballot would have to move for the same elections; "
                                        "above 1.0 because this kernel provably never touches those bytes (it splits %.0f nodes per "
                                        "election where a full expansion visits %.0f), not a bandwidth claim"
                                        % (st["items"] / per_el, denominators.get("kernel_counters", {}).get("items_full", float("nan")))}
            out["limiter"] = ("instruction issue and latency of dependent f32/f64 arithmetic in the Gamma / binomial samplers, "
                              "not HBM (profiles/r01i_ncu_lane_kernel_c3.txt, profiles/r01e_ncu_deferred_kernel_c3.txt)")
        return out

    rl = roofline(main)
    rl["denominators"] = denominators
    line = {
        "metric": METRIC, "value": value, "unit": "elections/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 Gamma variates + f64 binomial rejection; u32 counts/tallies", "data": "synthetic",
        "config": {"workload": describe(name, cfg, epg), "global_elections_per_step": epg * world,
                   "kernel": main["kernel"],
                   "cache": "L2 flushed between timed steps (a 512 MB buffer is zeroed outside the per-step event pairs; "
                            "%.2f ms for the %d steps with the flushes included); per-election scratch %.1f GB reserved, "
                            "nothing is reused between elections" % (main["elapsed_ms_with_flush"], args.steps, stats["scratch_bytes"] / 1e9),
                   "parallelism": "elections sharded over %d GPU(s), tree replicated, one all-reduce of %d win counts" % (world, nc),
                   "seed": SEED},
        "e2e": e2e, "gpu_launches": int(stats["launches"]) * args.steps,
        "arena_reruns_last_step": int(stats.get("arena_runs", 0)),
        "roofline": rl, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    if other is not None:
        key = "materialised" if other["kernel"] == "simulate_kernel" else "deferred"
        line[key] = {"kernel": other["kernel"], "value": other["value"], "unit": "elections/s", "steps": other["steps"],
                     "elections_per_gpu_per_step": other["elections_per_gpu"], "roofline": roofline(other),
                     "note": "same inputs and seed; outcomes are bit-identical to the headline kernel's"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
