#!/usr/bin/env python
"""bench.py — simulated IRV elections / second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3] [--impl ours|reference]

The hot path is Dirichlet-tree posterior draw -> IRV -> win counts (R_tree.cpp:177-257). A *job* is one call of it
over `--elections-per-gpu` elections of the named configuration on every GPU (weak scaling; the default C3 job is
BASELINE's 100 000 elections, one kernel launch). A *step* is R such jobs back to back on disjoint global election
ranges followed by one all-reduce of the <= 32 win counts (NCCL); R is fixed during warm-up so that a step lasts
about `--step-ms` (a single C3 job is ~1 ms, far too short to time robustly), identical on every rank.

 value      whole-job elections/s, tree resident in HBM, win counts left on the device; CUDA events on the launch
            stream around every step, L2 flushed between steps, max over ranks. median / min / max per step beside it.
 e2e        the same jobs through the host-buffer C-ABI call (dtree_sample_posterior) with the device mirror
            invalidated before every call: H2D of the ballot log + tree build + kernel + D2H of the result block.
 identity   win counts (and their SHA-256) of ONE fixed job — C3's 100 000 elections, seed BENCHSEEDx, global ids
            0..99 999 — sharded over the N ranks: identical for every N (north_star correctness part 3).
 strong     that same fixed job timed: BASELINE's "100k elections sharded over 8 GPUs" (strong scaling).
 workloads  (N = 1) the other BASELINE configurations and a close-race C3 shape, each with the kernel the library
            chose, node splits per election and its roofline on its own and on SURVEY 8(d)'s bytes.
 roofline   the headline kernel against measured HBM bandwidth (bytes defined in roofline()).
 materialised
            the same inputs through mode 0 (every ballot materialised: SURVEY 8(d)'s bytes), rated on those bytes.
 cpu_baseline / --impl reference
            the reference's own C++ (oracle/_ref, else the port) on all host cores, bounded sample.
"""
import argparse
import hashlib
import json
import math
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
DEFAULT_EPG = {"C1": 4_000_000, "C2": 1_000_000, "C3": 100_000, "C4": 9_472, "C5": 200_000}  # C4: 4 waves of 2 368 warps
KERNELS = {0: "simulate_kernel", 1: "deferred_kernel", 2: "lane_kernel"}


def describe(name, cfg):
    """The workload, in words; both arms of the bench print exactly this string."""
    obs = cfg["observed"]
    how = ("Plackett-Luce gamma=%g seed %d" % (obs[1], obs[3])) if obs[0] == "pl" else "Wakehurst sample seed %d" % obs[2]
    return ("%s: %d candidates, %d observed ballots (%s), n_ballots=%d, min_depth=%d, max_depth=%d, a0=%g, vd=%s"
            % (name, cfg["n_candidates"], obs[2] if obs[0] == "pl" else obs[1], how, cfg["n_ballots"], cfg["min_depth"],
               cfg["max_depth"], cfg["a0"], cfg["vd"]))


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
        "config": {"workload": describe(name, cfg), "elections_per_step": per_step,
                   "note": "CPU only; n_gpus is the launch shape, rank 0 does all the work"},
        "cpu_baseline": {"value": value, "unit": "elections/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "elections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def spread(ms):
    return {"median": float(np.median(ms)), "min": float(np.min(ms)), "max": float(np.max(ms))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--gamma", type=float, default=None, help="override the Plackett-Luce gamma of the observed ballots")
    ap.add_argument("--elections-per-gpu", type=int, default=0)
    ap.add_argument("--step-ms", type=float, default=100.0, help="a step repeats the job until it lasts about this long")
    ap.add_argument("--repeats", type=int, default=0, help="jobs per step (0: from --step-ms)")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--urn-max", type=int, default=-1)
    ap.add_argument("--tune", action="append", default=[], help="key=value for dtree_set_tuning (repeatable)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-mode", action="store_true", help="headline kernel only: no other-kernel, strong, workloads legs")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--mode", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="2/1: deferred expansion with a thread/warp per election, 0: materialising kernel, -1: the library's choice")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    pkg = load_pkg()
    W = pkg.workloads
    name = args.config
    cfg = dict(W.CONFIGS[name])
    if args.gamma is not None:
        cfg["observed"] = ("pl", args.gamma) + tuple(cfg["observed"][2:])
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
    L = N.lib()
    seed = SEED.encode()
    stream = torch.cuda.current_stream(dev)
    # Larger than the 126 MB L2: zeroed between timed steps so that no step starts with the previous one's lines.
    l2_flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        v = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def make_tree(wname, wcfg, ballots=None):
        nc_ = wcfg["n_candidates"]
        if ballots is None:
            rec = wcfg["observed"]
            ballots = W.plackett_luce_ballots(nc_, rec[1], rec[2], rec[3]) if rec[0] == "pl" else W.observed_ballots(wname)
        tr = pkg.dirichlet_tree(W.candidate_names(nc_), wcfg["min_depth"], wcfg["max_depth"], wcfg["a0"], wcfg["vd"], device=local)
        tr.update_indices(*ballots)
        if args.block_threads or args.blocks_per_sm:
            tr.tune(block_threads=args.block_threads, blocks_per_sm=args.blocks_per_sm)
        if args.urn_max >= 0:
            tr.tune(urn_max=args.urn_max)
        for kv in args.tune:
            k, v = kv.split("=")
            tr.tune(**{k: int(v)})
        N.check(L.dtree_sync_device(tr._h))
        return tr, ballots

    def measure(t, wcfg, mode, n_el, steps, warmup, offset, step_ms, repeats=0):
        """`steps` timed steps of R jobs of n_el elections per GPU, tree resident, counts left on the device."""
        nc_ = wcfg["n_candidates"]
        wins = torch.zeros(nc_, dtype=torch.int64, device=dev)
        total = torch.zeros(nc_, dtype=torch.int64, device=dev)
        t.tune(mode=mode)
        job = [0]  # global job counter: job k covers ids [offset + k*n_el*world, +n_el*world); this rank's contiguous share

        def launch():
            first = offset + (job[0] * world + rank) * n_el
            job[0] += 1
            N.check(L.dtree_sample_posterior_device(t._h, first, n_el, wcfg["n_ballots"], 1, 0, seed, len(seed),
                                                    wins.data_ptr(), stream.cuda_stream))

        def finish():
            N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))

        # warm-up: finish() lets the library resize its per-election scratch from what the elections used, so the
        # (re)allocations happen here and not inside the timed region
        for _ in range(max(3, warmup)):
            launch()
            finish()
        wins.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(3):
            launch()
        e1.record(stream)
        finish()
        job_ms = allmax(e0.elapsed_time(e1) / 3.0)
        R = repeats if repeats > 0 else int(max(1, min(512, math.ceil(step_ms / max(job_ms, 1e-3)))))
        if world > 1:
            dist.all_reduce(wins, op=dist.ReduceOp.SUM)
        total.add_(wins)  # first use of these torch ops loads their kernels: not inside the timed region
        l2_flush.zero_()
        total.zero_()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_ev, kern_ev = [], []
        ev0.record(stream)
        for i in range(steps):
            l2_flush.zero_()  # between timed steps, not inside one
            s0, k1, s1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            wins.zero_()
            s0.record(stream)
            for _ in range(R):
                launch()
            k1.record(stream)
            if world > 1:
                dist.all_reduce(wins, op=dist.ReduceOp.SUM)
            total.add_(wins)
            s1.record(stream)
            step_ev.append((s0, s1))
            kern_ev.append((s0, k1))
        ev1.record(stream)
        barrier()
        finish()
        st = t.last_stats()
        step_ms_l = [a.elapsed_time(b) for a, b in step_ev]
        kern_ms_l = [a.elapsed_time(b) / R for a, b in kern_ev]
        ms = allmax(sum(step_ms_l))
        ms_with_flush = allmax(ev0.elapsed_time(ev1))
        assert int(total.sum().item()) == n_el * world * steps * R, (int(total.sum().item()), n_el, world, steps, R)
        per_el = max(st["elections"], 1)
        return {"value": n_el * world * steps * R / (ms / 1e3), "elapsed_ms": ms, "elapsed_ms_with_flush": ms_with_flush,
                "steps": steps, "jobs_per_step": R, "elections_per_gpu": n_el,
                "kernel_ms": float(np.mean(kern_ms_l)), "kernel_ms_job": spread(kern_ms_l),
                "step_ms": spread(step_ms_l), "step_ms_steps": [round(x, 3) for x in step_ms_l],
                "after_kernels_ms": spread([k[1].elapsed_time(s_[1]) for k, s_ in zip(kern_ev, step_ev)]),
                "stats": st, "splits_per_election": st["items"] / per_el,
                "kernel": KERNELS[st["mode"]], "launches_per_job": st["launches"]}

    nc = cfg["n_candidates"]
    epg = args.elections_per_gpu or DEFAULT_EPG[name]
    t, (prefs, offsets) = make_tree(name, cfg)

    sampler = ClockSampler(local) if rank == 0 else None
    main = measure(t, cfg, args.mode, epg, args.steps, args.warmup, 0, args.step_ms, args.repeats)
    stats = main["stats"]
    # the other kernel on the same inputs (fewer elections when it is the slower, materialising one)
    other = None
    if not args.single_mode:
        om = 0 if stats["mode"] >= 1 else 1
        o_epg = max(296, epg // 8) if om == 0 else epg
        if om == 0 and o_epg > 1776:
            o_epg -= o_epg % 1776  # whole waves of the three phase kernels (12 elections per SM)
        # on a handle of its own, destroyed afterwards: the materialising path reserves tens of GB of per-election
        # scratch, which would otherwise stay with `t` and shrink what the later legs may plan with
        t_other, _ = make_tree(name, cfg, (prefs, offsets))
        other = measure(t_other, cfg, om, o_epg, max(2, min(args.steps, 3)), 3, 1 << 40, 0.0, 1)
        del t_other
        torch.cuda.empty_cache()
    t.tune(mode=args.mode)

    # ---- identity + strong scaling: ONE fixed job sharded over the ranks ----------------------------------------------
    identity = strong = None
    if not args.single_mode:
        n_job = cfg["n_elections"] if name != "C4" else 8 * DEFAULT_EPG["C4"]
        per = -(-n_job // world)
        lo = min(n_job, rank * per)
        hi = min(n_job, lo + per)
        wins = torch.zeros(nc, dtype=torch.int64, device=dev)

        def fixed_job():
            wins.zero_()
            if hi > lo:
                N.check(L.dtree_sample_posterior_device(t._h, lo, hi - lo, cfg["n_ballots"], 1, 0, seed, len(seed),
                                                        wins.data_ptr(), stream.cuda_stream))
            if world > 1:
                dist.all_reduce(wins, op=dist.ReduceOp.SUM)

        for _ in range(3):
            fixed_job()
            N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        barrier()
        reps = max(5, args.steps)
        evs = []
        for _ in range(reps):
            l2_flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fixed_job()
            b.record(stream)
            evs.append((a, b))
        barrier()
        N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        job_ms = [x.elapsed_time(y) for x, y in evs]
        tot_ms = allmax(sum(job_ms))
        vec = [int(x) for x in wins.cpu().tolist()]
        assert sum(vec) == n_job
        identity = {"job": "%s, %d elections, seed %s, global ids 0..%d, sharded over %d rank(s)" % (name, n_job, SEED, n_job - 1, world),
                    "win_counts": vec, "sha256": hashlib.sha256(np.asarray(vec, dtype=np.uint64).tobytes()).hexdigest()}
        strong = {"value": n_job * reps / (tot_ms / 1e3), "unit": "elections/s", "scaling": "strong",
                  "elections_total": n_job, "elections_per_gpu": per, "jobs_timed": reps, "ms_per_job": spread(job_ms),
                  "kernel": KERNELS[t.last_stats()["mode"]],
                  "note": "BASELINE config as stated: the fixed job split over the ranks (R_tree.cpp:200-207 partition), "
                          "one kernel launch per rank + one all-reduce per job"}

    # ---- e2e: host buffers through the C ABI, mirror invalidated so that the log upload + tree build is inside ------
    def e2e_call(first):
        N.check(L.dtree_invalidate_device(t._h))
        return t.sample_posterior_counts(epg, cfg["n_ballots"], seed=SEED, first_election=first)

    e2e_base = 1 << 41
    for i in range(3):  # untimed warm-up of this path too (it lets the library settle its scratch sizing)
        e2e_call(e2e_base + (i * world + rank) * epg)
    tw = time.perf_counter()
    for i in range(3, 6):
        e2e_call(e2e_base + (i * world + rank) * epg)
    call_ms = allmax((time.perf_counter() - tw) / 3.0 * 1e3)
    Re = int(max(1, min(256, math.ceil(args.step_ms / max(call_ms, 1e-3)))))
    host_us = {"mirror_us": 0, "plan_us": 0, "launch_us": 0, "finish_us": 0}
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    e2e_ms = []
    k = 6
    for i in range(args.steps):
        ts = time.perf_counter()
        acc = np.zeros(nc, dtype=np.int64)
        for _ in range(Re):
            w = e2e_call(e2e_base + (k * world + rank) * epg)
            k += 1
            acc += w.astype(np.int64)
            st_ = t.last_stats()
            h2d += st_["h2d_bytes"]
            d2h += st_["d2h_bytes"]
            for kk in host_us:
                host_us[kk] += st_[kk]
        if world > 1:
            wt = torch.from_numpy(acc).to(dev)
            dist.all_reduce(wt, op=dist.ReduceOp.SUM)
            acc = wt.cpu().numpy()
        e2e_ms.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    n_calls = args.steps * Re
    e2e = {"value": epg * world * n_calls / e2e_s, "unit": "elections/s", "h2d_bytes_per_step": int(h2d // args.steps),
           "d2h_bytes_per_step": int(d2h // args.steps), "steps": args.steps, "calls_per_step": Re,
           "ms_per_call": 1e3 * e2e_s / n_calls, "step_ms": spread(e2e_ms),
           "host_us_per_call": {kk: round(v / n_calls, 1) for kk, v in host_us.items()},
           "note": "every call: dtree_invalidate_device + dtree_sample_posterior with host buffers (ballot log re-uploaded, "
                   "tree rebuilt on the device, kernel, one read-back of the result block)"}
    clocks = sampler.stop() if sampler else None

    # ---- the other configurations (N = 1 only: they do not depend on the rank count) -----------------------------
    workloads = None
    if world == 1 and not args.single_mode and not args.no_workloads and name == "C3" and args.gamma is None:
        workloads = []
        c3close = dict(W.CONFIGS["C3"])
        c3close["observed"] = ("pl", 0.05) + tuple(c3close["observed"][2:])
        todo = [("C1", dict(W.CONFIGS["C1"]), 4_000_000, 592), ("C2", dict(W.CONFIGS["C2"]), 1_000_000, 592),
                ("C4", dict(W.CONFIGS["C4"]), DEFAULT_EPG["C4"], 148), ("C5", dict(W.CONFIGS["C5"]), 200_000, 592),
                ("C3 shape, gamma=0.05 (close race)", c3close, 20_000, 296)]
        for wname, wcfg, w_epg, mat_el in todo:
            wt_, _ = make_tree(wname.split(" ")[0], wcfg)
            m = measure(wt_, wcfg, -1, w_epg, 3, 3, 0, 150.0)
            mat = measure(wt_, wcfg, 0, mat_el, 2, 3, 1 << 40, 0.0, 1)  # the materialising kernel: 8(d)'s S, E, R
            workloads.append(workload_entry(wname, wcfg, m, mat, wt_))
            del wt_
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline denominators from the CPU oracle's counters on the same inputs -----------------
    mst = stats
    for cand in (main, other):
        if cand is not None and cand["kernel"] == "simulate_kernel":
            mst = cand["stats"]
    n_obs_entries = len(t.observed_indices()[2])
    S_k = mst["slots"] / max(mst["elections"], 1)
    E_k = mst["entries"] / max(mst["elections"], 1) + n_obs_entries
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

    rl = roofline(main, b_alg, nc, name if args.gamma is None else None)
    rl["denominators"] = denominators
    line = {
        "metric": METRIC, "value": main["value"], "unit": "elections/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["elapsed_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 Gamma variates + f64 binomial rejection; u32 counts/tallies", "data": "synthetic",
        "config": {"workload": describe(name, cfg), "elections_per_gpu_per_job": epg, "jobs_per_step": main["jobs_per_step"],
                   "global_elections_per_step": epg * world * main["jobs_per_step"], "kernel": main["kernel"],
                   "step": "R = %d jobs of %d elections per GPU back to back on disjoint global election ids (one kernel "
                           "launch + replay launch each), then one all-reduce of %d win counts; R fixed in warm-up so that "
                           "a step lasts ~%g ms" % (main["jobs_per_step"], epg, nc, args.step_ms),
                   "cache": "L2 flushed between timed steps (a 512 MB buffer is zeroed outside the per-step event pairs; "
                            "%.1f ms for the %d steps with the flushes included); inside a step the %.1f MB tree stays "
                            "L2-resident from job to job as it would in any long run, the per-election scratch (%.2f GB "
                            "reserved) is written before it is read"
                            % (main["elapsed_ms_with_flush"], args.steps, t_tree_mb(t), stats["scratch_bytes"] / 1e9),
                   "parallelism": "elections sharded over %d GPU(s), tree replicated, one all-reduce of %d win counts per step" % (world, nc),
                   "seed": SEED},
        "step_ms": main["step_ms"], "step_ms_steps": main["step_ms_steps"], "kernel_ms_per_job": main["kernel_ms_job"],
        "e2e": e2e, "gpu_launches": int(main["launches_per_job"]) * main["jobs_per_step"] * args.steps,
        "arena_reruns_last_job": int(stats.get("arena_runs", 0)), "lane_retries_last_job": int(stats.get("lane_retries", 0)),
        "splits_per_election": main["splits_per_election"],
        "identity": identity, "strong": strong,
        "roofline": rl, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    if other is not None:
        key = "materialised" if other["kernel"] == "simulate_kernel" else "deferred"
        line[key] = {"kernel": other["kernel"], "value": other["value"], "unit": "elections/s", "steps": other["steps"],
                     "elections_per_gpu_per_step": other["elections_per_gpu"],
                     "launches_per_job": int(other["launches_per_job"]),
                     "roofline": roofline(other, b_alg, nc, name if args.gamma is None else None),
                     "note": "same inputs and seed; outcomes are bit-identical to the headline kernel's. mode 0 runs a job of "
                             "more than two machine-fulls of elections as waves of three launches (sim_big_kernel, "
                             "sim_walk_kernel, sim_irv_kernel: DESIGN.md 5.2); kernel_ms is the whole job"}
    if workloads is not None:
        line["workloads"] = workloads
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def t_tree_mb(t):
    """Size of the device mirror of the tree (CSR arrays), MB."""
    import ctypes as C
    from elections_dtree_b200 import _native as N
    total = 0
    for which in range(5):
        need = C.c_uint64(0)
        N.check(t._L.dtree_debug_device_image(t._h, which, None, 0, C.byref(need)))
        total += need.value
    return total / 1e6


_PEAK = None


def hbm_peak():
    global _PEAK
    if _PEAK is None:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        _PEAK = (float(peaks.get("hbm_gbs", 6650.0)),
                 "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")
    return _PEAK


def roofline(m, b_alg, nc, traffic_config):
    """Algorithmic bytes per job / CUDA-event duration of the job, against measured HBM copy bandwidth.
    simulate_kernel moves SURVEY 8(d)'s bytes (B_alg = 4S + 32E + 16R + 8nc: every sampled ballot is written, read and
    re-read). The deferred kernels never create most of them, so they are rated on the bytes THEIR algorithm must
    move — 4 B per outcome slot read + every 32 B node record written to or read back from the per-election arena
    (counted by the kernel) + 8nc, DESIGN.md 5.1 — with the 8(d) rate beside it (`survey_8d`), which exceeds the HBM
    peak exactly when the kernel provably does not touch those bytes."""
    peak, peak_src = hbm_peak()
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        traffic_tab = {}
    n_el, secs = m["elections_per_gpu"], m["kernel_ms"] / 1e3
    st = m["stats"]
    per_el = max(st["elections"], 1)
    if m["kernel"] == "simulate_kernel":
        bytes_el = b_alg
    else:
        bytes_el = 4.0 * st["slots"] / per_el + 32.0 * (st["entries"] + st["irv_rereads"]) / per_el + 8 * nc
    ach = bytes_el * n_el / secs / 1e9
    tr = traffic_tab.get(m["kernel"], {})
    same = traffic_config is not None and tr.get("config") == traffic_config
    out = {"kernel": m["kernel"], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
           "traffic": tr["bytes_per_election"] * n_el if same else None,
           "traffic_source": tr.get("capture") if same else None,
           "peak_source": peak_src, "algorithmic_bytes_per_election": bytes_el,
           "kernel_ms_per_launch": m["kernel_ms"], "kernel_ms_per_job": m["kernel_ms_job"],
           "elections_per_launch": n_el, "splits_per_election": m["splits_per_election"]}
    if m["kernel"] != "simulate_kernel":
        eq = b_alg * n_el / secs / 1e9
        out["survey_8d"] = {"bytes_per_election": b_alg, "equivalent_GBps": eq, "equivalent_frac": eq / peak,
                            "note": "what a kernel that materialises every ballot would have to move for the same elections; "
                                    "above 1.0 means this kernel provably never touches those bytes, not a bandwidth claim"}
        out["limiter"] = ("instruction issue and latency of dependent arithmetic in the Gamma / binomial samplers and of dependent "
                          "loads from the per-election arenas, not HBM (profiles/)")
    return out


def workload_entry(wname, wcfg, m, mat, tree):
    """One row of `workloads`: the library's own choice of kernel on this configuration, rated on its own bytes and on
    SURVEY 8(d)'s (S, E, R from the materialising kernel's counters on the same inputs)."""
    nc = wcfg["n_candidates"]
    ms = mat["stats"]
    per = max(ms["elections"], 1)
    n_obs_entries = len(tree.observed_indices()[2])
    S, E, R = ms["slots"] / per, ms["entries"] / per + n_obs_entries, ms["irv_rereads"] / per
    b_alg = 4 * S + 32 * E + 16 * R + 8 * nc
    return {"workload": describe(wname, wcfg), "elections_per_gpu_per_job": m["elections_per_gpu"],
            "jobs_per_step": m["jobs_per_step"], "steps": m["steps"], "value": m["value"], "unit": "elections/s",
            "kernel": m["kernel"], "splits_per_election": m["splits_per_election"], "step_ms": m["step_ms"],
            "arena_reruns_last_job": m["stats"]["arena_runs"], "lane_retries_last_job": m["stats"]["lane_retries"],
            "scratch_gb": m["stats"]["scratch_bytes"] / 1e9,
            "roofline": roofline(m, b_alg, nc, None),
            "materialised": {"kernel": mat["kernel"], "value": mat["value"], "elections": mat["elections_per_gpu"],
                             "splits_per_election": mat["splits_per_election"], "S": S, "E": E, "R": R,
                             "roofline": roofline(mat, b_alg, nc, None)}}


if __name__ == "__main__":
    sys.exit(main())
