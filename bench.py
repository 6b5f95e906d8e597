#!/usr/bin/env python
"""bench.py -- DRRT hot-path benchmark (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the batched range-ball query (BASELINE config 2a:
1M nodes in (x,y,theta) with ghost wraparound, 64K queries per GPU at the RRTx
shrinking radius, gamma = 100) over one batch of synthetic queries.  Queries
shard across ranks with no data-path collective (weak scaling); the node store
is replicated by an NCCL broadcast of the insert batch before the timed region.
One JSON line is printed by rank 0.  `extra` carries the sparse-radius range
batch (2b), k-nearest, nearest, straight edges and the config-4 replanning run.

The line also carries (inside the objects the driver keeps): config 3 -- the 10 M-edge
Dubins collision batch -- with its own roofline (against the FP64 peak measured by
benchmarks/fp64_peak.py), CPU baseline and end-to-end figure through drrt_edgecheck_batch;
the narrow-row / id-only / fused-Extend end-to-end figures; config 1 (the planner iteration
over the reference's own objects, CPU vs shim); and, at N > 1, a parity record (every rank's
rows equal the single-GPU answer) and the strong-scaling figure of the one-process multi-GPU
C ABI (drrt_comm_*).

--impl reference times the reference's own CPU implementation of the same step:
oracle/_ref/drrt_ref, the reference's unmodified kdtree.cpp / ghostpoint.cpp / jlist.cpp
compiled against an Eigen stand-in (oracle/Makefile), one process per host core, on a bounded
query sample; the oracle port when that binary is absent.
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
        emit(line)
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
    emit(line)


def ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, parsed from the
    newest committed `ncu --set full` summary under profiles/ (None when there is none)"""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s*.txt" % kernel_prefix))):
        try:
            txt = open(path).read()
        except OSError:
            continue
        rd = re.search(r"dram__bytes_read\.sum\s+(?:\S+\s+)?([0-9.eE+]+)\s*(GB|MB|KB|B|Gbyte|Mbyte|Kbyte|byte)", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+(?:\S+\s+)?([0-9.eE+]+)\s*(GB|MB|KB|B|Gbyte|Mbyte|Kbyte|byte)", txt)
        if rd and wr:
            mul = {"GB": 1e9, "MB": 1e6, "KB": 1e3, "B": 1.0, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            best = (float(rd.group(1)) * mul[rd.group(2)] + float(wr.group(1)) * mul[wr.group(2)],
                    os.path.relpath(path, ROOT))
    return best


def edge_fp64_ops():
    """FP64 thread-level operations (DADD + DMUL + 2 x DFMA) of one config-3 step, counted by ncu
    (profiles/r02_edge_fp64_ops.json, the sum over the edge-check kernels); None when absent"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_edge_fp64_ops.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


def rows_checksum(torch, tree, res):
    """position-sensitive checksum of a range result (counts, ids, distance bit patterns), on the
    device: equal checksums <=> equal rows for all practical purposes"""
    res = tree.range_pack()
    off, ids, dist = tree.range_result_tensors(res)
    counts = tree.range_result_counts(res).to(torch.int64)
    t = int(res.total)
    w = (torch.arange(t, device=ids.device, dtype=torch.int64) % 65521) + 1
    wq = (torch.arange(counts.numel(), device=ids.device, dtype=torch.int64) % 65521) + 1
    return torch.stack([(counts * wq).sum(), (ids[:t].to(torch.int64) * w).sum(),
                        (dist[:t].view(torch.int64) % 1000003 * w).sum(),
                        torch.tensor(t, device=ids.device, dtype=torch.int64)])


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print to file descriptor 1 (NCCL's
    version banner, for one) is sent to stderr; emit() writes to the saved descriptor"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global N_NODES
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip config 3 / kNN / Extend / in-loop legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
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
    lib = capi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- replicated node store: rank 0's insert batch is broadcast over NCCL
    nodes_h = synth.nodes(N_NODES) if (rank == 0 or not args.no_extra) else None
    nodes_t = torch.empty((N_NODES, 3), dtype=torch.float64, device=dev)
    if rank == 0:
        nodes_t.copy_(torch.from_numpy(nodes_h))
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
        return rank_max(sum(a.elapsed_time(b) for a, b in evs))

    def host_timed(step, steps, warm=2):
        """wall time per call of a HOST entry point (copies inside), max over ranks"""
        for _ in range(warm):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        return rank_max(time.perf_counter() - t0) / steps

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

    # ---- N > 1: every rank's rows equal the single-GPU answer (checksums of counts / ids /
    # distance bits; rank 0 recomputes every shard on its own GPU, outside the timed region)
    parity = None
    if world > 1:
        mine = rows_checksum(torch, tree, state["res"])
        allsum = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsum, mine)
        if rank == 0:
            ok = True
            for r in range(world):
                qr = torch.from_numpy(np.ascontiguousarray(q_all[r * N_QUERIES:(r + 1) * N_QUERIES])).to(dev)
                want = rows_checksum(torch, tree, tree.KDFindWithinRange_dev(r2a, qr))
                ok = ok and bool(torch.equal(want.cpu(), allsum[r].cpu()))
            parity = {"ranks_equal_single_gpu": ok, "checked": "counts, ids and distance bit patterns of all "
                      "%d x %d config-2a queries: each rank's shard against rank 0's own answer" % (world, N_QUERIES)}
            if not ok:
                raise SystemExit("bench.py: a rank's rows differ from the single-GPU answer")
        barrier()

    # ---- end to end through the host C-ABI calls: pinned buffers, copies inside
    q_pin = torch.from_numpy(q_host).pin_memory()
    counts_pin = torch.empty(N_QUERIES, dtype=torch.int32).pin_memory()
    ids_pin = torch.empty(hits_2a + 1024, dtype=torch.int32).pin_memory()
    dist_pin = torch.empty(hits_2a + 1024, dtype=torch.float64).pin_memory()
    f32_pin = torch.empty(hits_2a + 1024, dtype=torch.float32).pin_memory()
    off_pin = torch.empty(N_QUERIES + 1, dtype=torch.int64).pin_memory()
    qp = C.cast(q_pin.data_ptr(), capi.c_dp)

    def into_step(kind):
        fn = lib.drrt_range_batch_into_f32 if kind == "f32" else lib.drrt_range_batch_into
        dp = {"f64": C.cast(dist_pin.data_ptr(), capi.c_dp), "f32": C.cast(f32_pin.data_ptr(), capi.c_fp),
              "ids": None}[kind]

        def f():
            total = C.c_int64()
            capi.check(fn(tree._h, N_QUERIES, qp, None, r2a, ids_pin.numel(), C.cast(off_pin.data_ptr(), capi.c_i64p),
                          C.cast(ids_pin.data_ptr(), capi.c_ip), dp, C.byref(total)))
            assert total.value == hits_2a
        return f

    def e2e_two_call_step():
        total = C.c_int64()
        capi.check(lib.drrt_range_batch(tree._h, N_QUERIES, qp, None, r2a, C.cast(counts_pin.data_ptr(), capi.c_ip),
                                        C.byref(total)))
        capi.check(lib.drrt_range_fetch(tree._h, C.cast(ids_pin.data_ptr(), capi.c_ip),
                                        C.cast(dist_pin.data_ptr(), capi.c_dp)))

    e2e_steps = max(2, min(args.steps, 5))
    launches0 = capi.launches()
    sec = host_timed(into_step("f64"), e2e_steps)
    e2e = {"value": world * N_QUERIES / sec, "unit": UNIT, "h2d_bytes_per_step": 24 * N_QUERIES,
           "d2h_bytes_per_step": 8 * (N_QUERIES + 1) + 12 * hits_2a,
           "call": "drrt_range_batch_into (one call, pinned host buffers, bit-exact double keys: 12 B per hit; "
                   "16384-query sub-batches, row copies overlap the next sub-batch)",
           "ms_per_step": 1e3 * sec,
           "gpu_launches_per_step": (capi.launches() - launches0) // (e2e_steps + 2)}
    sec = host_timed(into_step("f32"), e2e_steps)
    e2e["rows_f32"] = {"value": world * N_QUERIES / sec, "unit": UNIT, "ms_per_step": 1e3 * sec,
                       "d2h_bytes_per_step": 8 * (N_QUERIES + 1) + 8 * hits_2a,
                       "call": "drrt_range_batch_into_f32: ids bit-exact, distances rounded to float "
                               "(north star: 1e-5 relative) -- 8 B per hit"}
    sec = host_timed(into_step("ids"), e2e_steps)
    e2e["ids_only"] = {"value": world * N_QUERIES / sec, "unit": UNIT, "ms_per_step": 1e3 * sec,
                       "d2h_bytes_per_step": 8 * (N_QUERIES + 1) + 4 * hits_2a,
                       "call": "drrt_range_batch_into, dist == NULL: the neighbour id sets alone -- 4 B per hit"}

    roof_edge = None
    cpu_edge = None
    config5_summary = None
    extra = {}
    if not args.no_extra:
        sec = host_timed(e2e_two_call_step, e2e_steps)
        extra["e2e_two_call"] = {"queries_per_s": world * N_QUERIES / sec,
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

        # ---- config 3: 10 M Dubins edges = the query<->hit pairs of 2b, vs 256 polygons
        off_h, ids_h, _ = tree.KDFindWithinRange(r2b, q_host)
        s_h, e_h = synth.edges_from_neighbours(nodes_h, q_host, off_h, ids_h, N_EDGES)
        # the id form needs a node at both ends ((query -> hit) pairs have a free pose at one end): it is
        # timed on hit -> previous hit of the same 2b row, edges of the same length scale
        prev = np.arange(len(ids_h), dtype=np.int64) - 1      # the previous hit of the SAME row (cyclic)
        nz = np.diff(off_h) > 0
        prev[off_h[:-1][nz]] = off_h[1:][nz] - 1
        ids_a = np.ascontiguousarray(np.resize(ids_h, N_EDGES).astype(np.int32))
        ids_b = np.ascontiguousarray(np.resize(ids_h[prev], N_EDGES).astype(np.int32))
        del off_h, ids_h, prev, nz
        start = torch.from_numpy(s_h).to(dev)
        end = torch.from_numpy(e_h).to(dev)
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
        ms_edge = timed(step_edges(drrt_b200.EDGE_DUBINS), esteps, 3) / esteps
        coll = float(verdict.float().mean().item())
        fp64 = None
        try:
            from benchmarks import fp64_peak
            fp64 = fp64_peak.measure(local)
        except Exception as ex:  # the microbenchmark needs nvcc once; say so instead of guessing a peak
            fp64 = {"error": str(ex)[:200]}
        ops = edge_fp64_ops()
        roof_edge = {"bound": "fp64", "kernel": "dubins_rank + dubins_solve_bin x7 + dubins_walk (config 3: 10 M Dubins "
                     "edges vs 256 polygons, %.1f %% collide)" % (100 * coll),
                     "ms_per_step": ms_edge, "edges_per_s": world * N_EDGES / (ms_edge * 1e-3),
                     "hbm_frac": 57.0 * N_EDGES / (ms_edge * 1e-3) / 1e9 / peak, "bytes_per_edge": 57,
                     "fp64_peak": fp64}
        if ops and fp64 and "dfma_tflops" in fp64:
            # thread-level FP64 operations of one step (ncu count, profiles/) over the measured step time,
            # against the measured FP64 pipe: DADD / DMUL count 1, DFMA 2
            flop = float(ops["fp64_flop_per_step"])
            roof_edge.update({"achieved": flop / (ms_edge * 1e-3) / 1e12, "peak": fp64["dfma_tflops"],
                              "unit": "TFLOP/s", "frac": flop / (ms_edge * 1e-3) / 1e12 / fp64["dfma_tflops"],
                              # the library never fuses (bit parity): its ceiling is the instruction rate
                              "frac_of_unfused_peak": flop / (ms_edge * 1e-3) / 1e12 / fp64["dmul_dadd_tflops"],
                              "fp64_ops_source": ops.get("source")})
        ms = timed(step_edges(drrt_b200.EDGE_STRAIGHT), esteps, 3) / esteps
        extra["straight_edge_check"] = {"edges_per_s": world * N_EDGES / (ms * 1e-3), "ms_per_step": ms,
                                        "collision_rate": float(verdict.float().mean().item())}
        del start, end, verdict, length
        # end to end: host poses in (48 B per edge), verdicts + lengths out (9 B per edge)
        s_pin = torch.from_numpy(s_h).pin_memory()
        e_pin = torch.from_numpy(e_h).pin_memory()
        v_pin = torch.empty(N_EDGES, dtype=torch.uint8).pin_memory()
        l_pin = torch.empty(N_EDGES, dtype=torch.float64).pin_memory()
        a_pin = torch.from_numpy(ids_a).pin_memory()
        b_pin = torch.from_numpy(ids_b).pin_memory()

        def e2e_edges():
            capi.check(lib.drrt_edgecheck_batch(tree._h, N_EDGES, C.cast(s_pin.data_ptr(), capi.c_dp),
                                                C.cast(e_pin.data_ptr(), capi.c_dp), drrt_b200.EDGE_DUBINS,
                                                synth.ROBOT_RADIUS, synth.R_MIN, 0,
                                                C.cast(v_pin.data_ptr(), capi.c_u8p),
                                                C.cast(l_pin.data_ptr(), capi.c_dp)))

        def e2e_edges_ids():
            capi.check(lib.drrt_edgecheck_ids_batch(tree._h, N_EDGES, C.cast(a_pin.data_ptr(), capi.c_ip),
                                                    C.cast(b_pin.data_ptr(), capi.c_ip), drrt_b200.EDGE_DUBINS,
                                                    synth.ROBOT_RADIUS, synth.R_MIN, 0,
                                                    C.cast(v_pin.data_ptr(), capi.c_u8p),
                                                    C.cast(l_pin.data_ptr(), capi.c_dp)))

        sec = host_timed(e2e_edges, esteps)
        e2e["edge_check"] = {"value": world * N_EDGES / sec, "unit": "edges/s", "ms_per_step": 1e3 * sec,
                             "h2d_bytes_per_step": 48 * N_EDGES, "d2h_bytes_per_step": 9 * N_EDGES,
                             "collision_rate": float(v_pin.float().mean().item()),
                             "call": "drrt_edgecheck_batch (config 3, pinned host poses; 2 Mi-edge chunks: poses up, "
                                     "solve + walk, verdicts + lengths down on three streams)"}
        sec = host_timed(e2e_edges_ids, esteps)
        e2e["edge_check_ids"] = {"value": world * N_EDGES / sec, "unit": "edges/s", "ms_per_step": 1e3 * sec,
                                 "h2d_bytes_per_step": 8 * N_EDGES, "d2h_bytes_per_step": 9 * N_EDGES,
                                 "call": "drrt_edgecheck_ids_batch: both ends named by node id (8 B per edge up); "
                                         "node -> node edges between hits of the same 2b lists"}
        del s_pin, e_pin, a_pin, b_pin
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
        del ids_k, dist_k, ids_1, dist_1

        # ---- fused Extend (range + both Dubins edges per neighbour + best parent) on the 2b batch
        lmc_h = np.random.default_rng(0xD2270021).uniform(0.0, 80.0, N_NODES)
        lmc_t = torch.from_numpy(lmc_h).to(dev)
        bp = torch.empty(N_QUERIES, dtype=torch.int32, device=dev)
        bc = torch.empty(N_QUERIES, dtype=torch.float64, device=dev)

        def step_extend():
            state["ext"] = tree.Extend_dev(r2b, q_dev, lmc=lmc_t, best_parent=bp, best_cost=bc,
                                           robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN)

        ms_ext = timed(step_extend, esteps, 3) / esteps
        n_pairs = int(state["ext"][0].total)
        extra["extend_2b"] = {"new_nodes_per_s": world * N_QUERIES / (ms_ext * 1e-3), "ms_per_step": ms_ext,
                              "edge_checks_per_step": 2 * n_pairs,
                              "edge_checks_per_s": world * 2 * n_pairs / (ms_ext * 1e-3),
                              "with_parent": float((bp >= 0).float().mean().item())}
        del lmc_t
        # end to end: host in = the new nodes; host out = neighbour count, best parent, cost -- what
        # Extend / FindBestParent consume (drrt.cpp:194-362); rows and verdicts stay on the device,
        # rrt_LMC_ lives in the device mirror and takes a 4096-entry delta update per step
        tree.WriteLMC(0, lmc_h)
        bp_pin = torch.empty(N_QUERIES, dtype=torch.int32).pin_memory()
        bc_pin = torch.empty(N_QUERIES, dtype=torch.float64).pin_memory()
        ch_ids = np.ascontiguousarray(np.arange(0, N_NODES, max(1, N_NODES // 4096), dtype=np.int32)[:4096])
        ch_val = np.ascontiguousarray(lmc_h[ch_ids])

        def e2e_extend():
            total = C.c_int64()
            capi.check(lib.drrt_lmc_scatter(tree._h, len(ch_ids), ch_ids.ctypes.data_as(capi.c_ip),
                                            ch_val.ctypes.data_as(capi.c_dp)))
            capi.check(lib.drrt_extend_batch(tree._h, N_QUERIES, qp, None, r2b, synth.ROBOT_RADIUS, synth.R_MIN, 0,
                                             None, C.cast(counts_pin.data_ptr(), capi.c_ip), C.byref(total),
                                             C.cast(bp_pin.data_ptr(), capi.c_ip),
                                             C.cast(bc_pin.data_ptr(), capi.c_dp)))
            assert total.value == n_pairs

        sec = host_timed(e2e_extend, esteps)
        assert torch.equal(bp_pin, bp.cpu()) and torch.equal(bc_pin, bc.cpu())
        e2e["extend"] = {"value": world * N_QUERIES / sec, "unit": "new nodes/s", "ms_per_step": 1e3 * sec,
                         "device_ms_per_step": ms_ext, "over_device_time": 1e3 * sec / ms_ext,
                         "edge_checks_per_s": world * 2 * n_pairs / sec,
                         "h2d_bytes_per_step": 24 * N_QUERIES + 12 * len(ch_ids), "d2h_bytes_per_step": 16 * N_QUERIES,
                         "call": "drrt_lmc_scatter (4096 changed rrt_LMC_ entries) + drrt_extend_batch against the "
                                 "device mirror: counts, best parent and cost come back; %d neighbour rows and "
                                 "%d edge verdicts never cross PCIe" % (n_pairs, 2 * n_pairs)}
        del bp, bc

        # ---- single-call latency: the planner's own nq == 1 call (drrt.cpp:213-216) through the
        # C ABI, at 10 K nodes (its RRTx radius) and on the 1 M-node store (2b radius)
        small = drrt_b200.KDTree(device=local)
        small.KDInsert(synth.nodes(10_000, 0xD2270031))
        r10k = synth.shrinking_radius(10_000, synth.GAMMA_REF)
        one_off = np.zeros(2, dtype=np.int64)
        lat = {}
        for name, tr, rad in (("nodes_10k", small, r10k), ("nodes_1m", tree, r2b)):
            qs = synth.queries(512, 0xD2270032)
            ts = []
            nhit = 0
            for k in range(len(qs)):
                t0 = time.perf_counter()
                total = C.c_int64()
                capi.check(lib.drrt_range_batch_into(tr._h, 1, qs[k].ctypes.data_as(capi.c_dp), None, rad,
                                                     ids_pin.numel(), one_off.ctypes.data_as(capi.c_i64p),
                                                     C.cast(ids_pin.data_ptr(), capi.c_ip),
                                                     C.cast(dist_pin.data_ptr(), capi.c_dp), C.byref(total)))
                ts.append(time.perf_counter() - t0)
                nhit += total.value
            ts = np.sort(np.array(ts[64:]))
            lat[name] = {"radius": rad, "hits_per_query": nhit / len(qs), "us_median": 1e6 * float(np.median(ts)),
                         "us_p95": 1e6 * float(ts[int(0.95 * len(ts))])}
            ids1, d1 = tr.KDFindNearest(qs[:1])
            t0 = time.perf_counter()
            for k in range(64, 320):
                tr.KDFindNearest(qs[k:k + 1])
            lat[name]["nearest_us_mean"] = 1e6 * (time.perf_counter() - t0) / 256
        e2e["single_call"] = dict(lat, call="drrt_range_batch_into(nq = 1) / drrt_nearest_batch(nq = 1) via ctypes, "
                                            "wall clock per call including the copies")
        small.close()

        # ---- config 4: replanning scenario (2 M nodes grown by 4096-node batches), one store per rank
        from benchmarks import replanning
        extra["config4_replanning"] = replanning.run(2_000_000, 4096, device=local,
                                                      parity=(rank == 0 and world == 1 and not args.no_cpu))

        # ---- N > 1: BASELINE config 5 -- an 8 M-node store replicated by NCCL broadcast of 4096-node
        # insert batches, 64 K queries per GPU at its RRTx radius; every rank's rows against rank 0's
        # own answer (the store is twice the L2: the HBM-bound point of the range kernel)
        if world > 1:
            extra["config5_8m_nodes"] = config5(torch, dist, drrt_b200, synth, dev, local, rank, world, timed, state,
                                                peak, max(2, min(args.steps, 5)))
            config5_summary = extra["config5_8m_nodes"]

        # ---- N > 1: strong scaling of the ONE-PROCESS multi-GPU C ABI (drrt_comm_*): rank 0's process
        # drives all N GPUs (the reference planner is one process); the other ranks wait at the barrier
        if world > 1:
            import datetime
            barrier()    # every rank's GPU is idle from here on
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                try:
                    e2e["strong_scaling"] = comm_strong_scaling(world, nodes_h, q_all[:N_QUERIES], r2a, hits_2a,
                                                                s_h, e_h)
                finally:
                    store.set("drrt_strong_scaling_done", "1")
            else:
                # a CPU-side wait: an NCCL barrier would park a spinning kernel on the very GPUs rank 0's
                # process is timing
                store.wait(["drrt_strong_scaling_done"], datetime.timedelta(seconds=1200))
            barrier()
        del s_h, e_h

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_tree, cpu = cpu_range_baseline(r2a, 0)

        def cpu_tree_for_single_calls():
            return cpu_tree, None
        if N_NODES == 1_000_000:
            # beside the port: the reference's own compiled sources on the same queries (what
            # `--impl reference` times at length), a short sample
            refbin = reference_binary_rate(r2a, 8, 0, 1)
            if refbin is not None:
                cpu["reference_binary"] = {k: refbin[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                                  "hits_per_query")}
        if roof_edge is not None:
            # the oracle's CalculateTrajectory + ExplicitEdgeCheck on a bounded sample of config-3 edges
            from oracle import pyoracle as po
            off_h, ids_h, _ = tree.KDFindWithinRange(r2b, q_host[:2048])
            s2, e2 = synth.edges_from_neighbours(nodes_h, q_host[:2048], off_h, ids_h, 300_000)
            robs = po.RefObstacles(**synth.obstacles())
            t0 = time.perf_counter()
            robs.edge_check_batch(s2, e2, nthreads=0)
            dt = time.perf_counter() - t0
            cpu_edge = {"value": len(s2) / dt, "unit": "edges/s", "cores": po.num_threads(), "kind": "port",
                        "sample": "%d config-3 edges, oracle Dubins solve + polyline + ExplicitEdgeCheck2D with "
                                  "first-hit exit, %.1f s" % (len(s2), dt)}
            refbin = reference_binary_edge_rate(s2, e2, synth.obstacles(), 2000)
            if refbin is not None:
                cpu_edge["reference_binary"] = refbin
            cpu["edge_check"] = cpu_edge
            # one query per call on the CPU (the oracle's tree walk, one thread), beside e2e.single_call
            if "single_call" in e2e:
                from oracle import pyoracle as po
                small = po.RefTree()
                small.insert(synth.nodes(10_000, 0xD2270031))
                qs = synth.queries(512, 0xD2270032)
                big, _ = cpu_tree_for_single_calls()
                sc = {}
                for name, tr, rad in (("nodes_10k", small, synth.shrinking_radius(10_000, synth.GAMMA_REF)),
                                      ("nodes_1m", big, r2b)):
                    ts = []
                    for k in range(64, 320):
                        t0 = time.perf_counter()
                        tr.range_one(qs[k], rad)
                        ts.append(time.perf_counter() - t0)
                    sc[name] = {"us_median": 1e6 * float(np.median(ts)), "radius": rad}
                sc["call"] = "the oracle's KDFindWithinRange, one query per call, one thread (ctypes)"
                cpu["single_call"] = sc
            # config 1: the planner iteration over the reference's OWN objects, all-CPU reference vs
            # the same driver over the shim + this library (oracle/_ref/drrt_ref vs drrt_shim)
            inloop = config1_inloop()
            if inloop is not None:
                cpu["inloop"] = inloop["cpu"]
                e2e["inloop"] = inloop["gpu"]

    if rank == 0:
        traffic = ncu_traffic("range_cta_kernel") if N_NODES == 1_000_000 else None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on the
                    # default workload, parsed from the committed ncu --set full summary (null when
                    # there is none, or for other store sizes)
                    "traffic": traffic[0] if traffic else None,
                    "traffic_source": traffic[1] if traffic else None,
                    "traffic_unit": "bytes/launch",
                    "peak_kind": peak_kind,
                    "kernel": "range_cta_kernel<R2S1> (CTA per query)",
                    "algorithmic_bytes_per_launch": alg_bytes}
        if roof_edge is not None:
            roofline["edge_check"] = roof_edge
        config = {"workload": "config2a batched range-ball: %d nodes (x,y,theta), ghost "
                              "wraparound, 64K queries/GPU, r=%.6f (RRTx shrinking radius, "
                              "gamma=100)" % (N_NODES, r2a),
                  "nodes": N_NODES, "queries_per_gpu": N_QUERIES,
                  "hits_per_query": hits_2a / N_QUERIES,
                  "l2": "flushed between timed steps (512 MiB memset, untimed)",
                  "result": "rows in HBM, sorted by id, addressed by offsets+counts "
                            "(slices of 8 consecutive queries)",
                  "store": "replicated per GPU; NCCL broadcast of the insert batch"}
        if parity is not None:
            config["parity_across_gpus"] = parity
        if config5_summary is not None:
            config["config5_8m_nodes"] = {k: config5_summary[k] for k in
                                          ("queries_per_s", "ms_per_step", "roofline_frac", "ranks_equal_single_gpu")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "extra": extra}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def config5(torch, dist, drrt_b200, synth, dev, local, rank, world, timed, state, peak, steps):
    """BASELINE config 5 on this run's GPUs (called at N > 1)"""
    n5, batch = 8_000_000, 4096
    tree = drrt_b200.KDTree(device=local, capacity=n5)
    src = torch.from_numpy(synth.nodes(n5, synth.SEED_NODES_8GPU)).to(dev) if rank == 0 else None
    buf = torch.empty((batch, 3), dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    for lo in range(0, n5, batch):
        m = min(batch, n5 - lo)
        if rank == 0:
            buf[:m].copy_(src[lo:lo + m])
        dist.broadcast(buf, src=0)                    # the insert batch reaches every replica over NVLink
        tree.KDInsert_dev(buf[:m].contiguous())
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    del src
    r5 = synth.shrinking_radius(n5, synth.GAMMA_REF)
    q_all = synth.queries(N_QUERIES * world, synth.SEED_QUERIES + 5)
    qd = torch.from_numpy(np.ascontiguousarray(q_all[rank * N_QUERIES:(rank + 1) * N_QUERIES])).to(dev)

    def f():
        state["res5"] = tree.KDFindWithinRange_dev(r5, qd)

    ms = timed(f, steps, 3) / steps
    hits = int(state["res5"].total)
    mine = rows_checksum(torch, tree, state["res5"])
    allsum = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allsum, mine)
    ok = None
    if rank == 0:
        ok = True
        for r in range(world):
            qr = torch.from_numpy(np.ascontiguousarray(q_all[r * N_QUERIES:(r + 1) * N_QUERIES])).to(dev)
            want = rows_checksum(torch, tree, tree.KDFindWithinRange_dev(r5, qr))
            ok = ok and bool(torch.equal(want.cpu(), allsum[r].cpu()))
        if not ok:
            raise SystemExit("bench.py: config 5 -- a rank's rows differ from the single-GPU answer")
    dist.barrier()
    out = {"nodes": n5, "insert_batches": (n5 + batch - 1) // batch, "build_s": t_build, "radius": r5,
           "hits_per_query": hits / N_QUERIES, "ms_per_step": ms, "queries_per_s": world * N_QUERIES / (ms * 1e-3),
           "roofline_frac": (32.0 * N_QUERIES + 36.0 * hits) / (ms * 1e-3) / 1e9 / peak,
           "ranks_equal_single_gpu": ok,
           "what": "8 M-node store replicated on %d GPUs by NCCL broadcast of 4096-node insert batches, 64 K "
                   "queries per GPU; counts / ids / distance bits of every rank's shard equal rank 0's own answer" % world}
    tree.close()
    torch.cuda.empty_cache()
    return out


def config1_inloop(iters=1200):
    """BASELINE config 1: ref_driver.cpp's planner loop (sample -> KDFindNearest -> Saturate ->
    ExplicitNodeCheck -> KDFindWithinRange -> 2|N| edges -> KDInsert) over the reference's own
    KDTree / JList / Edge objects: all CPU (drrt_ref) vs the shim + libdrrt_b200.so (drrt_shim)."""
    from oracle import pyref
    from drrt_b200 import synth
    if not (os.path.exists(pyref.BIN) and os.path.exists(pyref.SHIM_BIN)):
        return None
    o = synth.obstacles(24, 0xD2270012, lo=8.0, hi=42.0)
    try:
        cpu = pyref.loop_time(o, iters, binary=pyref.BIN)
        per_edge = pyref.loop_time(o, iters, mode=0, binary=pyref.SHIM_BIN)
        batched = pyref.loop_time(o, iters, mode=1, binary=pyref.SHIM_BIN)
    except Exception:
        return None
    same = all(cpu[k] == per_edge[k] == batched[k] for k in ("nodes", "hits", "edges", "blocked", "key_sum"))
    what = ("%d iterations of the smalltest.cpp loop (delta 5, gamma 100, 24 polygons): %d nodes inserted, "
            "%d neighbours, %d Dubins edges" % (iters, cpu["nodes"], cpu["hits"], cpu["edges"]))
    return {"cpu": {"value": iters / cpu["seconds"], "unit": "iterations/s", "cores": 1, "kind": "reference",
                    "edges_per_s": cpu["edges"] / cpu["seconds"],
                    "sample": what + "; the reference's kdtree.cpp + CalculateTrajectory + analytic checker, one "
                                     "thread (its tree calls are serialised under tree_mutex_)"},
            "gpu": {"value": iters / batched["seconds"], "unit": "iterations/s",
                    "per_edge_calls": iters / per_edge["seconds"],
                    "edges_per_s": batched["edges"] / batched["seconds"],
                    "same_nodes_and_neighbour_keys_as_cpu": same,
                    "call": "the same driver over drrt_b200/shim: every KDTree call is one C-ABI call per query; "
                            "`value` checks the 2|N| edges of an iteration in one ExplicitEdgeCheckBatchGpu call "
                            "(the INTEGRATION.md patch), `per_edge_calls` one ExplicitEdgeCheckGpu call per edge "
                            "(the planner unchanged)"}}


def comm_strong_scaling(n_gpus, nodes_h, q_h, radius, hits, s_h, e_h):
    """fixed work -- 64 K config-2a queries and 10 M config-3 edges -- split over n_gpus GPUs by ONE
    host process through drrt_comm_*; wall clock per call, host buffers, copies inside"""
    from drrt_b200 import synth
    from drrt_b200.comm import CommKDTree, pinned_empty
    c = CommKDTree(n_gpus, capacity=len(nodes_h))
    t0 = time.perf_counter()
    c.KDInsert(nodes_h)
    t_ins = time.perf_counter() - t0
    c.SetObstacles(**synth.obstacles())
    nq = len(q_h)
    off, k0 = pinned_empty(nq + 1, np.int64)
    ids, k1 = pinned_empty(hits + 1024, np.int32)
    d64, k2 = pinned_empty(hits + 1024, np.float64)
    d32, k3 = pinned_empty(hits + 1024, np.float32)
    out = {"n_gpus": n_gpus, "insert_ms": 1e3 * t_ins, "nccl_broadcast_bytes": c.broadcast_bytes,
           "call": "one process, drrt_comm_range_batch_into[_f32] / drrt_comm_edgecheck_batch: contiguous shards, "
                   "every GPU copies its rows straight into the caller's pinned buffers; NCCL carries the insert "
                   "broadcast only (24 B x nodes x (N - 1))"}

    def wall(f, n=3):
        f()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        return (time.perf_counter() - t0) / n

    for name, dist_buf, per_hit in (("rows_f64", d64, 12), ("rows_f32", d32, 8), ("ids_only", None, 4)):
        sec = wall(lambda: c.KDFindWithinRangeInto(radius, q_h, off, ids, dist_buf))
        assert int(off[nq]) == hits
        out[name] = {"queries_per_s": nq / sec, "ms_per_step": 1e3 * sec, "d2h_bytes_per_step": per_hit * hits}
    ne = len(s_h)
    v, k4 = pinned_empty(ne, np.uint8)
    ln, k5 = pinned_empty(ne, np.float64)
    sp, k6 = pinned_empty(3 * ne, np.float64)
    ep, k7 = pinned_empty(3 * ne, np.float64)
    sp[:] = s_h.reshape(-1)
    ep[:] = e_h.reshape(-1)
    sec = wall(lambda: c.ExplicitEdgeCheck(sp.reshape(-1, 3), ep.reshape(-1, 3), robot_radius=synth.ROBOT_RADIUS,
                                           r_min=synth.R_MIN, verdict=v, length=ln), 2)
    out["edge_check"] = {"edges_per_s": ne / sec, "ms_per_step": 1e3 * sec, "collision_rate": float(v.mean()),
                         "h2d_bytes_per_step": 48 * ne, "d2h_bytes_per_step": 9 * ne}
    c.close()
    return out


if __name__ == "__main__":
    main()
