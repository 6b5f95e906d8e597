#!/usr/bin/env python
"""BASELINE config 4: largetest.cpp-scale replanning on one GPU.

Box [-50,50]^2 x [0,2pi') (src/largetest.cpp:253-254), delta = 10, gamma = 100.
The store grows to --nodes (default 2 M) by insert batches of 4096; after every
batch the batch's own nodes are range-queried at the RRTx radius and every
(new -> hit) and (hit -> new) Dubins edge is collision-checked against 256
polygons; every 64 batches 16 obstacles move by (+-1, +-1) and the
obstacle-conflict queries of FindPointsInConflictWithObstacle
(src/obstacle.cpp:540-610: radius robot_radius + O.radius + PI', centre theta = PI')
are re-issued.  Everything stays on the device; one JSON line is printed.

    python benchmarks/replanning.py [--nodes 2000000] [--batch 4096]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(nodes=2_000_000, batch=4096, device=0, parity=False):
    """grows one store on `device`; returns the result dict (torch's current stream is used).
    parity: at the final size, a 64-node sample of the last batch is answered by the CPU oracle
    too (its tree over the same nodes, the obstacle table after all moves): rows bit-equal,
    verdicts equal outside the 1e-6 band; the oracle's time on that sample is the CPU leg."""
    class args:  # noqa: N801
        pass
    args.nodes, args.batch = nodes, batch
    import torch
    import drrt_b200
    from drrt_b200 import capi, synth
    from drrt_b200.kdtree import KDTree

    dev = torch.device("cuda", device)
    if torch.cuda.current_stream().cuda_stream == 0:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # the C ABI treats stream 0 as "its own"
    rng = np.random.default_rng(0xD2270004)
    pos = np.empty((args.nodes, 3))
    pos[:, 0] = rng.uniform(-50, 50, args.nodes)
    pos[:, 1] = rng.uniform(-50, 50, args.nodes)
    pos[:, 2] = rng.uniform(0, synth.TWO_PI, args.nodes)
    pos_t = torch.from_numpy(pos).to(dev)
    obs = synth.obstacles(256, 0xD2270003, lo=-48.0, hi=48.0)
    tree = KDTree(device=device, capacity=args.nodes)
    tree.SetObstacles(**obs)
    delta, gamma = 10.0, 100.0

    n_batches = args.nodes // args.batch
    tot_q = tot_hits = tot_edges = tot_coll = tot_conf_q = tot_conf_hits = 0
    t_insert = t_range = t_edge = t_conf = 0.0
    ev = lambda: torch.cuda.Event(enable_timing=True)
    l0 = capi.launches()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for b in range(n_batches):
        lo, hi = b * args.batch, (b + 1) * args.batch
        batch = pos_t[lo:hi]
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        tree.KDInsert_dev(batch)
        e1.record()
        r = synth.shrinking_radius(hi, gamma, delta)
        res = tree.KDFindWithinRange_dev(r, batch)
        res = tree.range_pack()  # the edge list below is built from a plain CSR
        e2.record()
        off, ids, _ = KDTree.range_result_tensors(res)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(lo, hi, device=dev, dtype=torch.int32), counts)
        start = torch.cat([src, ids])
        end = torch.cat([ids, src])
        verdict = torch.empty(start.shape[0], dtype=torch.uint8, device=dev)
        length = torch.empty(start.shape[0], dtype=torch.float64, device=dev)
        tree.ExplicitEdgeCheck_dev(verdict, start_ids=start, end_ids=end, robot_radius=synth.ROBOT_RADIUS,
                                   r_min=synth.R_MIN, length_out=length)
        e3.record()
        torch.cuda.synchronize()
        if os.environ.get('REPLAN_DEBUG') and (b % 40 == 0 or e1.elapsed_time(e2) > 2.0):
            print(b, 'insert %.3f range %.3f edge %.3f edges %d' % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), start.shape[0]), file=sys.stderr, flush=True)
        t_insert += e0.elapsed_time(e1)
        t_range += e1.elapsed_time(e2)
        t_edge += e2.elapsed_time(e3)
        tot_q += args.batch
        tot_hits += int(res.total)
        tot_edges += int(start.shape[0])
        tot_coll += int(verdict.sum().item())
        if (b + 1) % 64 == 0:
            # 16 obstacles move; nodes in conflict with each are looked up again
            which = rng.choice(256, 16, replace=False)
            step = rng.choice([-1.0, 1.0], (16, 2))
            for k, o in enumerate(which):
                obs["origin"][o] += step[k]
                obs["verts"][obs["vert_off"][o]:obs["vert_off"][o + 1]] += step[k]
            # Obstacle::UpdateObstacles (obstacle.cpp:182-195): only the 16 moved entries cross the boundary
            tree.MoveObstacles(which, obs["origin"][which], translate_polygon=True)
            centres = np.concatenate([obs["origin"][which], np.full((16, 1), synth.PI)], 1)
            radii = synth.ROBOT_RADIUS + obs["radius"][which] + synth.PI
            c0, c1 = ev(), ev()
            c0.record()
            resc = tree.KDFindWithinRange_dev(torch.from_numpy(radii).to(dev), torch.from_numpy(centres).to(dev))
            c1.record()
            torch.cuda.synchronize()
            t_conf += c0.elapsed_time(c1)
            tot_conf_q += 16
            tot_conf_hits += int(resc.total)
    wall = time.perf_counter() - wall0
    line = {"workload": "config4 replanning: grow to %d nodes by %d-node insert batches; per batch range + "
                        "2*hits Dubins edge checks; 16 obstacles move every 64 batches" % (args.nodes, args.batch),
            "nodes": n_batches * args.batch, "batches": n_batches,
            "wall_s": wall, "gpu_ms": {"insert": t_insert, "range": t_range, "edge_check": t_edge,
                                       "conflict_range": t_conf},
            "inserts_per_s": n_batches * args.batch / (t_insert * 1e-3),
            "range_queries_per_s": tot_q / (t_range * 1e-3), "hits_per_query": tot_hits / tot_q,
            "edge_checks_per_s": tot_edges / (t_edge * 1e-3), "edges": tot_edges,
            "collision_rate": tot_coll / max(tot_edges, 1),
            "conflict_queries": tot_conf_q, "conflict_hits_per_query": tot_conf_hits / max(tot_conf_q, 1),
            "gpu_launches": capi.launches() - l0}
    if parity:
        line["parity_sample"], line["cpu_baseline"] = parity_sample(tree, pos, obs, n_batches * args.batch,
                                                                    synth.shrinking_radius(n_batches * args.batch, gamma, delta))
    tree.close()
    return line


def parity_sample(tree, pos, obs, n, radius, nq=64):
    """the last batch's first nq nodes against the CPU oracle at the final store size"""
    from oracle import pyoracle as po
    from drrt_b200 import synth
    ref = po.RefTree()
    t0 = time.perf_counter()
    ref.insert(pos[:n])
    t_build = time.perf_counter() - t0
    q = pos[n - nq:n]
    off, ids, dist = tree.KDFindWithinRange(radius, q)
    t0 = time.perf_counter()
    roff, rids, rdist = ref.range_batch(q, radius)
    t_range = time.perf_counter() - t0
    rows_equal = bool(np.array_equal(off, roff) and np.array_equal(ids, rids) and np.array_equal(dist, rdist))
    src = np.repeat(np.arange(n - nq, n), np.diff(off))
    s = np.concatenate([pos[src], pos[ids]])
    e = np.concatenate([pos[ids], pos[src]])
    v, ln = tree.ExplicitEdgeCheck(s, e, robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN)
    robs = po.RefObstacles(**obs)
    t0 = time.perf_counter()
    rv, rln = robs.edge_check_batch(s, e, robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN)
    t_edge = time.perf_counter() - t0
    diff = np.nonzero(v != rv)[0]
    _, _, clr = robs.edge_check_batch(s[diff], e[diff], robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN,
                                      want_clearance=True)
    # the sample's queries are nodes of the store, so every row holds the query itself: a node -> same
    # node edge is the degenerate case of CalculateTrajectory (equal poses: whether a turn of 0 or of a
    # full 2 pi comes out hangs on the last bit of an atan2, SURVEY 7 "hard parts") -- the planner never
    # builds one (the new node enters the tree after its range query, drrt.cpp:213-242); lengths are
    # compared on the other pairs
    self_pair = np.concatenate([src == ids, src == ids])
    ok = (rln < 1e11) & ~self_pair
    return ({"nodes": n, "queries": nq, "hits": int(len(ids)), "rows_bit_equal": rows_equal, "edges": int(len(v)),
             "verdicts_differ": int(len(diff)), "all_differences_within_1e-6_band": bool((clr < 1e-6).all()),
             "lengths_within_1e-9": bool(np.allclose(ln[ok], rln[ok], rtol=1e-9, atol=1e-9)),
             "self_pairs_left_out_of_the_length_comparison": int(self_pair.sum())},
            {"kind": "port", "cores": po.num_threads(), "range_queries_per_s": nq / t_range,
             "edge_checks_per_s": len(v) / t_edge,
             "sample": "%d queries + their %d Dubins edges at %d nodes (oracle tree build %.1f s not counted)"
                       % (nq, len(v), n, t_build)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    a = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    print(json.dumps(run(a.nodes, a.batch)), flush=True)


if __name__ == "__main__":
    main()
