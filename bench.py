#!/usr/bin/env python
"""bench.py -- DRRT hot-path benchmark (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the batched range-ball query (BASELINE config 2a:
1M nodes in (x,y,theta) with ghost wraparound, 64K queries per GPU at the RRTx
shrinking radius, gamma = 100) over one batch of synthetic queries.  Queries
shard across ranks with no data-path collective (weak scaling); the node store
is replicated by an NCCL broadcast of the insert batch before the timed region.
One JSON line is printed by rank 0.  `extra` carries the sparse-radius range
batch (2b), k-nearest, nearest and the 10M-edge Dubins collision batch
(config 3), each with its own roofline figures.

--impl reference times the reference's own CPU algorithm for the same step: the
oracle restatement of KDTree::KDFindWithinRange (the reference itself needs
Eigen/Bullet and cannot be built here), all host threads, on a
bounded query sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES = 1_000_000   # --nodes overrides (config 5 uses 8 M)
N_QUERIES = 65_536
N_EDGES = 10_000_000
K_NN = 16
METRIC = "range_ball_queries_per_s"
UNIT = "queries/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.max_sm, self.reasons = [], 0.0, set()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(float(out[0]))
                self.max_sm = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm or None,
                "reasons": sorted(self.reasons)}


def cpu_range_baseline(radius, nthreads, budget_s=20.0):
    """times the oracle's KDFindWithinRange on a bounded query sample"""
    from oracle import pyoracle as po
    from drrt_b200 import synth
    tree = po.RefTree()
    tree.insert(synth.nodes(N_NODES))
    q = synth.queries(N_QUERIES)
    n = 256
    t0 = time.perf_counter()
    tree.range_counts(q[:n], radius, nthreads=nthreads)
    dt = time.perf_counter() - t0
    n2 = int(min(N_QUERIES, max(n, n * budget_s / max(dt, 1e-6))))
    n2 = max(256, (n2 // 256) * 256)
    t0 = time.perf_counter()
    counts = tree.range_counts(q[:n2], radius, nthreads=nthreads)
    dt = time.perf_counter() - t0
    cores = nthreads if nthreads > 0 else po.num_threads()
    return tree, {"value": n2 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                  "sample": "%d of %d config-2a queries, one KDFindWithinRange each (oracle tree walk "
                            "incl. ghost pass), %d threads, %.1f s" % (n2, N_QUERIES, cores, dt),
                  "hits_per_query": float(counts.mean())}


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "drrt_ref")


def reference_binary_rate(radius, q_per_proc, warm, timed, procs=None):
    """Times the REFERENCE'S OWN kdtree.cpp / ghostpoint.cpp / jlist.cpp (oracle/_ref/drrt_ref, built
    by `make -C oracle ref` from the unmodified sources under /root/reference): one KDFindWithinRange
    + EmptyRangeList per query.  The reference serialises every tree call under tree_mutex_ and marks
    found nodes in_heap_, so one tree serves one thread: to use all host cores, one process per core
    builds its own 1 M-node tree (~1.4 GB, ~5 s) and answers its own slice of the queries; the rates
    of the concurrently running processes are added.  None when the binary is absent."""
    if not os.path.exists(REF_BIN):
        return None
    import tempfile
    import numpy as np
    from drrt_b200 import synth
    from oracle import pyref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = procs or cores
    try:
        import psutil
        procs = max(1, min(procs, int(psutil.virtual_memory().available // (2 << 30))))
    except Exception:
        pass
    pos = synth.nodes(N_NODES)
    q = synth.queries(N_QUERIES)
    q_per_proc = max(1, min(q_per_proc, N_QUERIES // procs))
    with tempfile.TemporaryDirectory() as tmp:
        jobs = []
        for p in range(procs):
            qs = np.ascontiguousarray(q[p * q_per_proc:(p + 1) * q_per_proc])
            fin, fout = os.path.join(tmp, "in%d.bin" % p), os.path.join(tmp, "out%d.bin" % p)
            with open(fin, "wb") as f:
                f.write(pyref.range_time_payload(pos, qs, radius, warm, timed))   # R2S1 metric, theta wraps
            jobs.append((fout, len(qs)))
        t0 = time.perf_counter()
        running = [subprocess.Popen([REF_BIN, "range_time", os.path.join(tmp, "in%d.bin" % p), jobs[p][0]],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                   for p in range(procs)]
        rcs = [r.wait() for r in running]
        wall = time.perf_counter() - t0
        if any(rcs):
            return None
        rate, hits, nq, longest = 0.0, 0, 0, 0.0
        for fout, n in jobs:
            h, secs = pyref.parse_time(open(fout, "rb").read(16))
            rate += n * timed / secs
            hits += h
            nq += n
            longest = max(longest, secs)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": "%d processes x %d config-2a queries x %d timed passes (+%d warm-up), each process the "
                      "reference's own KDFindWithinRange + EmptyRangeList on its own 1 M-node KDTree "
                      "(oracle/_ref/drrt_ref: kdtree.cpp, ghostpoint.cpp, jlist.cpp unmodified, Eigen "
                      "stand-in); slowest process %.1f s of queries, %.1f s wall with the tree builds"
                      % (procs, q_per_proc, timed, warm, longest, wall),
            "hits_per_query": hits / max(nq, 1), "queries_per_pass": nq,
            "ms_per_pass": 1e3 * longest / max(timed, 1)}


def reference_binary_edge_rate(s_h, e_h, obstacles, per_proc, procs=None):
    """The reference's own Edge::NewEdge + DubinsEdge::CalculateTrajectory + analytic ExplicitEdgeCheck per
    edge (oracle/_ref/drrt_ref edgecheck_time: dubinsedge.cpp, obstacle.cpp, distancefunctions.cpp
    unmodified), one process per core on its own slice of the edges.  None when the binary is absent."""
    if not os.path.exists(REF_BIN):
        return None
    import tempfile
    from drrt_b200 import synth
    from oracle import pyref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = procs or cores
    per_proc = max(1, min(per_proc, len(s_h) // procs))
    with tempfile.TemporaryDirectory() as tmp:
        outs = []
        for p in range(procs):
            lo, hi = p * per_proc, (p + 1) * per_proc
            fin, fout = os.path.join(tmp, "in%d.bin" % p), os.path.join(tmp, "out%d.bin" % p)
            with open(fin, "wb") as f:
                f.write(pyref.edge_time_payload(s_h[lo:hi], e_h[lo:hi], obstacles, synth.R_MIN, synth.ROBOT_RADIUS))
            outs.append(fout)
        running = [subprocess.Popen([REF_BIN, "edgecheck_time", os.path.join(tmp, "in%d.bin" % p), outs[p]],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                   for p in range(procs)]
        if any(r.wait() for r in running):
            return None
        rate, hits, longest = 0.0, 0, 0.0
        for fout in outs:
            h, secs = pyref.parse_time(open(fout, "rb").read(16))
            rate += per_proc / secs
            hits += h
            longest = max(longest, secs)
    return {"value": rate, "unit": "edges/s", "cores": procs, "kind": "reference",
            "sample": "%d processes x %d config-3 edges, the reference's own NewEdge + CalculateTrajectory + "
                      "ExplicitEdgeCheck2D loop per edge (oracle/_ref/drrt_ref, Eigen stand-in), slowest process "
                      "%.1f s" % (procs, per_proc, longest),
            "collision_rate": hits / float(procs * per_proc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from drrt_b200 import synth
    radius = synth.shrinking_radius(N_NODES, synth.GAMMA_REF)
    total_steps = max(1, args.steps + args.warmup)
    # (a) the reference's own compiled sources, when they were built (oracle/_ref travels with the
    # repo); ~0.36 s per query and process: size the per-step sample for ~100 s of queries
    per_proc = int(max(2, min(64, 100.0 / (0.4 * total_steps))))
    ref = reference_binary_rate(radius, per_proc, args.warmup, max(1, args.steps))
    if ref is not None:
        line = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ref["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config2a range-ball: 1M nodes, r=%.6f (gamma=100), %d-query sample "
                                       "per step" % (radius, ref["queries_per_pass"])},
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample", "hits_per_query")},
                "e2e": {"value": ref["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    # (b) the oracle port (plain-C restatement, bit-identical results), all host threads
    from oracle import pyoracle as po
    threads = po.num_threads()
    tree = po.RefTree()
    tree.insert(synth.nodes(N_NODES))
    q = synth.queries(N_QUERIES)
    # size the per-step sample so the whole run stays within a few minutes
    t0 = time.perf_counter()
    tree.range_counts(q[:256], radius)
    per_q = (time.perf_counter() - t0) / 256
    sample = int(min(N_QUERIES, max(256, 120.0 / max(per_q * total_steps, 1e-9))))
    sample = max(256, (sample // 256) * 256)
    for _ in range(args.warmup):
        tree.range_counts(q[:sample], radius)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tree.range_counts(q[:sample], radius)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2a range-ball: 1M nodes, r=%.6f (gamma=100), "
                                   "%d-query sample per step" % (radius, sample)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d config-2a queries per step x %d steps" % (sample, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    global N_NODES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the kNN / edge-check extras")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--nodes", type=int, default=N_NODES,
                    help="nodes in the replicated store (BASELINE config 5: 8000000)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    N_NODES = args.nodes

    if args.impl == "reference":
        run_reference(args)
        return

    # NCCL writes its version / debug lines to stdout by default: keep stdout for the one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import drrt_b200
    from drrt_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; drrt_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    capi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- replicated node store: rank 0's insert batch is broadcast over NCCL
    nodes_t = torch.empty((N_NODES, 3), dtype=torch.float64, device=dev)
    if rank == 0:
        nodes_t.copy_(torch.from_numpy(synth.nodes(N_NODES)))
    if world > 1:
        dist.broadcast(nodes_t, src=0)
    tree = drrt_b200.KDTree(device=local, capacity=N_NODES)
    tree.KDInsert_dev(nodes_t)
    torch.cuda.synchronize()
    # everything timed runs on one explicit stream: the C ABI treats a NULL
    # stream as "the handle's own", and CUDA events must sit on the launching stream
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    # ---- this rank's shard of the global query batch (contiguous index range)
    q_all = synth.queries(N_QUERIES * world)
    q_host = np.ascontiguousarray(q_all[rank * N_QUERIES:(rank + 1) * N_QUERIES])
    q_dev = torch.from_numpy(q_host).to(dev)
    r2a = synth.shrinking_radius(N_NODES, synth.GAMMA_REF)
    r2b = synth.shrinking_radius(N_NODES, synth.GAMMA_SPARSE)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    state = {}

    def timed(fn, steps, warmup):
        """device time of `steps` calls, L2 flushed (untimed) between them"""
        for _ in range(warmup):
            fn()
        barrier()
        state["launches0"] = capi.launches()
        evs = []
        for _ in range(steps):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        state["launches"] = capi.launches() - state["launches0"]
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_kind = measured_peak_gbs()

    def step_range(radius):
        def f():
            state["res"] = tree.KDFindWithinRange_dev(radius, q_dev)
        return f

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_range(r2a), args.steps, args.warmup)
    launches = state["launches"]
    clocks = sampler.result() if rank == 0 else None
    hits_2a = int(state["res"].total)
    ms_step = ms_total / args.steps
    value = world * N_QUERIES / (ms_step * 1e-3)
    alg_bytes = 32.0 * N_QUERIES + 36.0 * hits_2a
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9

    # ---- end to end through the host C-ABI call: pinned buffers, copies inside
    q_pin = torch.from_numpy(q_host).pin_memory()
    counts_pin = torch.empty(N_QUERIES, dtype=torch.int32).pin_memory()
    ids_pin = torch.empty(hits_2a + 1024, dtype=torch.int32).pin_memory()
    dist_pin = torch.empty(hits_2a + 1024, dtype=torch.float64).pin_memory()
    lib = capi.load()

    off_pin = torch.empty(N_QUERIES + 1, dtype=torch.int64).pin_memory()

    def e2e_step():
        # the one-call host query: rows of a finished sub-batch cross PCIe while the next
        # sub-batch is computed
        total = C.c_int64()
        capi.check(lib.drrt_range_batch_into(tree._h, N_QUERIES, C.cast(q_pin.data_ptr(), capi.c_dp),
                                             None, r2a, ids_pin.numel(),
                                             C.cast(off_pin.data_ptr(), capi.c_i64p),
                                             C.cast(ids_pin.data_ptr(), capi.c_ip),
                                             C.cast(dist_pin.data_ptr(), capi.c_dp), C.byref(total)))
        return total.value

    def e2e_two_call_step():
        total = C.c_int64()
        capi.check(lib.drrt_range_batch(tree._h, N_QUERIES, C.cast(q_pin.data_ptr(), capi.c_dp), None,
                                        r2a, C.cast(counts_pin.data_ptr(), capi.c_ip),
                                        C.byref(total)))
        capi.check(lib.drrt_range_fetch(tree._h, C.cast(ids_pin.data_ptr(), capi.c_ip),
                                        C.cast(dist_pin.data_ptr(), capi.c_dp)))
        return total.value

    def e2e_time(step):
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            assert step() == hits_2a
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * N_QUERIES * e2e_steps / float(dt.item())

    e2e_steps = max(2, min(args.steps, 5))
    launches0 = capi.launches()
    e2e_value = e2e_time(e2e_step)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * N_QUERIES,
           "d2h_bytes_per_step": 8 * (N_QUERIES + 1) + 12 * hits_2a,
           "call": "drrt_range_batch_into (one call, pinned host buffers; 16384-query sub-batches, "
                   "row copies overlap the next sub-batch)",
           "gpu_launches_per_step": (capi.launches() - launches0) // (e2e_steps + 2)}

    extra = {}
    if not args.no_extra:
        extra["e2e_two_call"] = {"queries_per_s": e2e_time(e2e_two_call_step),
                                 "call": "drrt_range_batch (counts) + drrt_range_fetch (rows): kernel, "
                                         "pack, then the copies"}
        # 2b: sparse radius (~76 hits/query)
        ms = timed(step_range(r2b), args.steps, 3) / args.steps
        res = state["res"]
        hits_2b = int(res.total)
        bytes_2b = 32.0 * N_QUERIES + 36.0 * hits_2b
        extra["range_2b"] = {"queries_per_s": world * N_QUERIES / (ms * 1e-3), "ms_per_step": ms,
                             "radius": r2b, "hits_per_query": hits_2b / N_QUERIES,
                             "roofline_frac": bytes_2b / (ms * 1e-3) / 1e9 / peak}
        # edges of config 3 from the 2b neighbour lists (built once, untimed)
        off_h, ids_h, _ = tree.KDFindWithinRange(r2b, q_host)
        s_h, e_h = synth.edges_from_neighbours(synth.nodes(N_NODES) if rank else nodes_t.cpu().numpy(),
                                               q_host, off_h, ids_h, N_EDGES)
        start = torch.from_numpy(s_h).to(dev)
        end = torch.from_numpy(e_h).to(dev)
        del s_h, e_h, off_h, ids_h
        o = synth.obstacles()
        tree.SetObstacles(**o)
        verdict = torch.empty(N_EDGES, dtype=torch.uint8, device=dev)
        length = torch.empty(N_EDGES, dtype=torch.float64, device=dev)

        def step_edges(kind):
            def f():
                tree.ExplicitEdgeCheck_dev(verdict, start=start, end=end, robot_radius=synth.ROBOT_RADIUS,
                                           r_min=synth.R_MIN, edge_kind=kind, length_out=length)
            return f

        esteps = max(2, min(args.steps, 5))
        ms = timed(step_edges(drrt_b200.EDGE_DUBINS), esteps, 3) / esteps
        extra["dubins_edge_check"] = {"edges_per_s": world * N_EDGES / (ms * 1e-3), "ms_per_step": ms,
                                      "collision_rate": float(verdict.float().mean().item()),
                                      "roofline_frac": 57.0 * N_EDGES / (ms * 1e-3) / 1e9 / peak,
                                      "bytes_per_edge": 57}
        ms = timed(step_edges(drrt_b200.EDGE_STRAIGHT), esteps, 3) / esteps
        extra["straight_edge_check"] = {"edges_per_s": world * N_EDGES / (ms * 1e-3),
                                        "ms_per_step": ms,
                                        "collision_rate": float(verdict.float().mean().item())}
        del start, end, verdict, length
        # kNN (k=16) and nearest on the same queries
        ids_k = torch.empty((N_QUERIES, K_NN), dtype=torch.int32, device=dev)
        dist_k = torch.empty((N_QUERIES, K_NN), dtype=torch.float64, device=dev)
        ms = timed(lambda: tree.KDFindKNearest_dev(K_NN, q_dev, ids_k, dist_k), args.steps, 3) / args.steps
        extra["knn_k16"] = {"queries_per_s": world * N_QUERIES / (ms * 1e-3), "ms_per_step": ms,
                            "roofline_frac": (24.0 + 36.0 * K_NN) * N_QUERIES / (ms * 1e-3) / 1e9 / peak}
        ids_1 = torch.empty(N_QUERIES, dtype=torch.int32, device=dev)
        dist_1 = torch.empty(N_QUERIES, dtype=torch.float64, device=dev)
        ms = timed(lambda: tree.KDFindKNearest_dev(1, q_dev, ids_1, dist_1), args.steps, 3) / args.steps
        extra["nearest"] = {"queries_per_s": world * N_QUERIES / (ms * 1e-3), "ms_per_step": ms,
                            "roofline_frac": 60.0 * N_QUERIES / (ms * 1e-3) / 1e9 / peak}

        # fused Extend (range + both Dubins edges per neighbour + best parent) on the 2b batch
        lmc_t = torch.rand(N_NODES, dtype=torch.float64, device=dev) * 80.0
        bp = torch.empty(N_QUERIES, dtype=torch.int32, device=dev)
        bc = torch.empty(N_QUERIES, dtype=torch.float64, device=dev)

        def step_extend():
            state["ext"] = tree.Extend_dev(r2b, q_dev, lmc=lmc_t, best_parent=bp, best_cost=bc,
                                           robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN)

        ms = timed(step_extend, esteps, 3) / esteps
        n_pairs = int(state["ext"][0].total)
        extra["extend_2b"] = {"new_nodes_per_s": world * N_QUERIES / (ms * 1e-3), "ms_per_step": ms,
                              "edge_checks_per_step": 2 * n_pairs,
                              "edge_checks_per_s": world * 2 * n_pairs / (ms * 1e-3),
                              "with_parent": float((bp >= 0).float().mean().item())}
        del lmc_t, bp, bc
        # config 4: replanning scenario (2 M nodes grown by 4096-node batches), one store per rank
        del ids_k, dist_k, ids_1, dist_1
        from benchmarks import replanning
        extra["config4_replanning"] = replanning.run(2_000_000, 4096, device=local)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        _, cpu = cpu_range_baseline(r2a, 0)
        if N_NODES == 1_000_000:
            # beside the port: the reference's own compiled sources on the same queries (what
            # `--impl reference` times at length), a short sample
            refbin = reference_binary_rate(r2a, 8, 0, 1)
            if refbin is not None:
                cpu["reference_binary"] = {k: refbin[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                                  "hits_per_query")}
        if "dubins_edge_check" in extra:
            # the oracle's CalculateTrajectory + ExplicitEdgeCheck on a bounded sample of config-3 edges
            from oracle import pyoracle as po
            off_h, ids_h, _ = tree.KDFindWithinRange(r2b, q_host[:2048])
            s_h, e_h = synth.edges_from_neighbours(synth.nodes(N_NODES), q_host[:2048], off_h, ids_h, 200_000)
            robs = po.RefObstacles(**synth.obstacles())
            t0 = time.perf_counter()
            robs.edge_check_batch(s_h, e_h, nthreads=0)
            dt = time.perf_counter() - t0
            extra["dubins_edge_check"]["cpu_baseline"] = {
                "value": len(s_h) / dt, "unit": "edges/s", "cores": po.num_threads(), "kind": "port",
                "sample": "%d config-3 edges, oracle Dubins solve + polyline + ExplicitEdgeCheck2D with "
                          "first-hit exit, %.1f s" % (len(s_h), dt)}
            refbin = reference_binary_edge_rate(s_h, e_h, synth.obstacles(), 2000)
            if refbin is not None:
                extra["dubins_edge_check"]["cpu_baseline"]["reference_binary"] = refbin

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "config2a batched range-ball: %d nodes (x,y,theta), ghost "
                                       "wraparound, 64K queries/GPU, r=%.6f (RRTx shrinking radius, "
                                       "gamma=100)" % (N_NODES, r2a),
                           "nodes": N_NODES, "queries_per_gpu": N_QUERIES,
                           "hits_per_query": hits_2a / N_QUERIES,
                           "l2": "flushed between timed steps (512 MiB memset, untimed)",
                           "result": "rows in HBM, sorted by id, addressed by offsets+counts "
                                     "(slices of 8 consecutive queries)",
                           "store": "replicated per GPU; NCCL broadcast of the insert batch"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel
                             # on the default workload (ncu --set full, profiles/r01v12_range_cta_kernel.txt);
                             # not captured for other store sizes
                             "traffic": 3.450e9 if N_NODES == 1_000_000 else None,
                             "traffic_unit": "bytes/launch",
                             "peak_kind": peak_kind,
                             "kernel": "range_cta_kernel<R2S1> (CTA per query)",
                             "algorithmic_bytes_per_launch": alg_bytes},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
