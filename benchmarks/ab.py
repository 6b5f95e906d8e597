#!/usr/bin/env python
"""A/B numbers of one build of the library (DRRT_B200_LIB selects it; default = the in-tree one):

    python benchmarks/ab.py [range] [knn] [single] [edges] [loop]

knn    : device time of k-nearest (k = 16) and nearest on the same queries, checksums
range  : device time of the headline step (config 2a, 64 K queries, L2 flushed), 2b, hit checksum
single : wall-clock per nq == 1 call of drrt_range_batch_into / drrt_nearest_batch / one-edge check
edges  : config 3 -- device time and the host call (pinned poses) of 10 M Dubins edges
loop   : the planner loop of config 1 over the shim (oracle/_ref/drrt_shim), both edge-call modes
"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import drrt_b200  # noqa: E402
from drrt_b200 import capi, synth  # noqa: E402

what = set(sys.argv[1:]) or {"range", "single", "edges"}
tag = os.environ.get("DRRT_B200_LIB", "default").split("/")[-2:]
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
lib = capi.load()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def dev_time(fn, steps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


nodes = synth.nodes()
q = synth.queries()
tree = drrt_b200.KDTree(device=0, capacity=len(nodes))
tree.KDInsert(nodes)
r2a = synth.shrinking_radius(len(nodes), synth.GAMMA_REF)
r2b = synth.shrinking_radius(len(nodes), synth.GAMMA_SPARSE)
qd = torch.from_numpy(q).to(dev)

if "range" in what:
    st = {}

    def f(r):
        def g():
            st["res"] = tree.KDFindWithinRange_dev(r, qd)
        return g
    for name, r in (("2a", r2a), ("2b", r2b)):
        mean, mn = dev_time(f(r))
        res = tree.range_pack()
        off, ids, dist = tree.range_result_tensors(res)
        chk = int(ids[:res.total].to(torch.int64).sum().item()) ^ int(dist[:res.total].view(torch.int64).sum().item())
        print(tag, "range", name, "ms mean %.4f min %.4f" % (mean, mn), "hits", int(res.total), "checksum", chk, flush=True)

if "knn" in what:
    K = 16
    kid = torch.empty((len(q), K), dtype=torch.int32, device=dev)
    kd = torch.empty((len(q), K), dtype=torch.float64, device=dev)
    for name, k in (("k16", K), ("nearest", 1)):
        mean, mn = dev_time(lambda: tree.KDFindKNearest_dev(k, qd, kid, kd))
        n = len(q) * k
        chk = int(kid.view(-1)[:n].to(torch.int64).sum().item()) ^ int(kd.view(-1)[:n].view(torch.int64).sum().item())
        print(tag, "knn", name, "ms mean %.4f min %.4f" % (mean, mn), "checksum", chk, flush=True)

if "single" in what:
    small = drrt_b200.KDTree(device=0)
    small.KDInsert(synth.nodes(10_000, 0xD2270031))
    small.SetObstacles(**synth.obstacles(24, 0xD2270012, lo=8.0, hi=42.0))
    ids = np.empty(1 << 16, dtype=np.int32)
    dist = np.empty(1 << 16, dtype=np.float64)
    off = np.zeros(2, dtype=np.int64)
    for name, tr, rad in (("10k", small, synth.shrinking_radius(10_000)), ("1m", tree, r2b)):
        qs = synth.queries(600, 0xD2270032)
        ts, tn, te = [], [], []
        for k in range(len(qs)):
            total = C.c_int64()
            t0 = time.perf_counter()
            capi.check(lib.drrt_range_batch_into(tr._h, 1, qs[k].ctypes.data_as(capi.c_dp), None, rad, len(ids),
                                                 off.ctypes.data_as(capi.c_i64p), ids.ctypes.data_as(capi.c_ip),
                                                 dist.ctypes.data_as(capi.c_dp), C.byref(total)))
            ts.append(time.perf_counter() - t0)
            nid, nd = np.empty(1, dtype=np.int32), np.empty(1, dtype=np.float64)
            t0 = time.perf_counter()
            capi.check(lib.drrt_nearest_batch(tr._h, 1, qs[k].ctypes.data_as(capi.c_dp), nid.ctypes.data_as(capi.c_ip),
                                              nd.ctypes.data_as(capi.c_dp)))
            tn.append(time.perf_counter() - t0)
        print(tag, "single", name, "range us median %.1f p95 %.1f | nearest us median %.1f"
              % (1e6 * np.median(ts[50:]), 1e6 * np.percentile(ts[50:], 95), 1e6 * np.median(tn[50:])), flush=True)
    s, e = synth.random_edges(600, 0xD2270033)
    v, ln = np.empty(1, dtype=np.uint8), np.empty(1, dtype=np.float64)
    te = []
    for k in range(len(s)):
        t0 = time.perf_counter()
        capi.check(lib.drrt_edgecheck_batch(small._h, 1, s[k].ctypes.data_as(capi.c_dp), e[k].ctypes.data_as(capi.c_dp),
                                            0, 0.5, 1.0, 0, v.ctypes.data_as(capi.c_u8p), ln.ctypes.data_as(capi.c_dp)))
        te.append(time.perf_counter() - t0)
    print(tag, "single one-edge check us median %.1f" % (1e6 * np.median(te[50:])), flush=True)

if "edges" in what:
    off, ids, _ = tree.KDFindWithinRange(r2b, q)
    s_h, e_h = synth.edges_from_neighbours(nodes, q, off, ids, 10_000_000)
    tree.SetObstacles(**synth.obstacles())
    n = len(s_h)
    start, end = torch.from_numpy(s_h).to(dev), torch.from_numpy(e_h).to(dev)
    verdict = torch.empty(n, dtype=torch.uint8, device=dev)
    length = torch.empty(n, dtype=torch.float64, device=dev)
    mean, mn = dev_time(lambda: tree.ExplicitEdgeCheck_dev(verdict, start=start, end=end, robot_radius=0.5, r_min=1.0,
                                                           length_out=length), steps=5)
    print(tag, "edges device ms mean %.3f min %.3f" % (mean, mn), "collide", int(verdict.sum().item()), flush=True)
    mean, mn = dev_time(lambda: tree.ExplicitEdgeCheck_dev(verdict, start=start, end=end, robot_radius=0.5, r_min=1.0,
                                                           edge_kind=1, length_out=length), steps=5)
    print(tag, "straight edges device ms mean %.3f min %.3f" % (mean, mn), "collide", int(verdict.sum().item()), flush=True)
    del start, end
    sp, ep = torch.from_numpy(s_h).pin_memory(), torch.from_numpy(e_h).pin_memory()
    vp, lp = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        capi.check(lib.drrt_edgecheck_batch(tree._h, n, C.cast(sp.data_ptr(), capi.c_dp), C.cast(ep.data_ptr(), capi.c_dp),
                                            0, 0.5, 1.0, 0, C.cast(vp.data_ptr(), capi.c_u8p), C.cast(lp.data_ptr(), capi.c_dp)))
        ts.append(time.perf_counter() - t0)
    print(tag, "edges host call ms min %.3f (%s)" % (1e3 * min(ts[2:]), ["%.1f" % (1e3 * t) for t in ts]),
          "collide", int(vp.sum().item()), flush=True)

if "loop" in what:
    from oracle import pyref
    o = synth.obstacles(24, 0xD2270012, lo=8.0, hi=42.0)
    for mode in (0, 1):
        r = pyref.loop_time(o, 1200, mode=mode, binary=pyref.SHIM_BIN)
        print(tag, "loop mode", mode, "iterations/s %.0f" % (1200 / r["seconds"]), r, flush=True)
