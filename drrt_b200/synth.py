"""Deterministic synthetic workloads of SURVEY.md section 8d (BASELINE.json configs).

Everything is numpy with fixed seeds so the GPU path, the CPU oracle and every
rank of a multi-GPU run see the same nodes, queries, obstacles and edges.
"""
import math

import numpy as np

from .capi import PI

TWO_PI = 2.0 * PI  # the reference's wrap point, 2.0*PI with PI = 3.1415926536

SEED_NODES = 0xD2270001
SEED_QUERIES = 0xD2270002
SEED_OBSTACLES = 0xD2270003
SEED_NODES_8GPU = 0xD2270005

DELTA = 5.0           # src/smalltest.cpp:215
GAMMA_REF = 100.0     # src/smalltest.cpp:216
GAMMA_SPARSE = 27.46  # SURVEY 8d config 2b (~76 hits/query at 1M nodes)
ROBOT_RADIUS = 0.5    # src/smalltest.cpp:199-ish parameter block
R_MIN = 1.0           # src/smalltest.cpp:198


def box_points(n, seed, lo=0.0, hi=50.0):
    """n i.i.d. uniform poses in [lo,hi)^2 x [0, 2*PI')."""
    rng = np.random.default_rng(seed)
    p = np.empty((n, 3), dtype=np.float64)
    p[:, 0] = rng.uniform(lo, hi, n)
    p[:, 1] = rng.uniform(lo, hi, n)
    p[:, 2] = rng.uniform(0.0, TWO_PI, n)
    return p


def nodes(n=1_000_000, seed=SEED_NODES):
    return box_points(n, seed)


def queries(n=65_536, seed=SEED_QUERIES):
    return box_points(n, seed)


def shrinking_radius(n_nodes, gamma=GAMMA_REF, delta=DELTA, dims=3):
    """RRTx ball radius min(delta, gamma*(ln(1+N)/N)^(1/d)) with the intended
    real-valued exponent (the reference's integer 1/d makes it
    min(delta, gamma); src/mainloop.cpp:64-67)."""
    return min(delta, gamma * (math.log(1.0 + n_nodes) / n_nodes) ** (1.0 / dims))


def obstacles(m=256, seed=SEED_OBSTACLES, lo=2.0, hi=48.0):
    """m kind-3 star-shaped polygons, world-frame vertices,
    origin_ = centre, radius_ = max vertex distance as obstacle.h:134-141."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(lo, hi, m)
    cy = rng.uniform(lo, hi, m)
    circ = rng.uniform(0.5, 2.0, m)
    nv = rng.integers(3, 9, m)
    vert_off = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(nv, out=vert_off[1:])
    verts = np.empty((int(vert_off[-1]), 2), dtype=np.float64)
    radius = np.empty(m, dtype=np.float64)
    for i in range(m):
        ang = np.sort(rng.uniform(0.0, 2.0 * math.pi, nv[i]))
        rad = circ[i] * rng.uniform(0.6, 1.0, nv[i])
        vx = cx[i] + rad * np.cos(ang)
        vy = cy[i] + rad * np.sin(ang)
        verts[vert_off[i]:vert_off[i + 1], 0] = vx
        verts[vert_off[i]:vert_off[i + 1], 1] = vy
        d = np.sqrt((vx - cx[i]) ** 2 + (vy - cy[i]) ** 2)
        radius[i] = d.max()
    return dict(kind=np.full(m, 3, dtype=np.int32), used=np.ones(m, dtype=np.uint8),
                life_span=np.full(m, 1000000000000.0), origin=np.stack([cx, cy], 1),
                radius=radius, vert_off=vert_off, verts=verts)


def edges_from_neighbours(node_pos, query_pos, offsets, ids, n_edges):
    """The planner's edge list: for each query in order,
    every (query -> hit) and (hit -> query) pair; truncated, or cycled, to
    exactly n_edges.  Returns (start[n,3], end[n,3])."""
    counts = np.diff(offsets)
    qidx = np.repeat(np.arange(len(counts)), counts)
    qs = query_pos[qidx]
    hs = node_pos[ids]
    start = np.empty((2 * len(ids), 3), dtype=np.float64)
    end = np.empty_like(start)
    start[0::2] = qs
    end[0::2] = hs
    start[1::2] = hs
    end[1::2] = qs
    if len(start) == 0:
        raise ValueError("no neighbours to build edges from")
    if len(start) >= n_edges:
        return np.ascontiguousarray(start[:n_edges]), np.ascontiguousarray(end[:n_edges])
    reps = -(-n_edges // len(start))
    return (np.ascontiguousarray(np.tile(start, (reps, 1))[:n_edges]),
            np.ascontiguousarray(np.tile(end, (reps, 1))[:n_edges]))


def random_edges(n, seed, max_len=0.66, lo=0.0, hi=50.0):
    """n edges whose end points lie within max_len (in x,y) of their starts."""
    rng = np.random.default_rng(seed)
    s = box_points(n, seed + 1, lo, hi)
    ang = rng.uniform(0, 2 * math.pi, n)
    ln = max_len * np.sqrt(rng.uniform(0, 1, n))
    e = np.empty_like(s)
    e[:, 0] = s[:, 0] + ln * np.cos(ang)
    e[:, 1] = s[:, 1] + ln * np.sin(ang)
    e[:, 2] = rng.uniform(0.0, TWO_PI, n)
    return s, e


def write_obstacle_file(path, table):
    """Writes a polygon table in the grammar of Obstacle::ReadObstaclesFromFile
    (src/obstacle.cpp:25-155): polygon count; per polygon a delimiter line, `ox,oy`, the number
    of moves (0 here: static), the vertex count and one `x,y` line per vertex.  Vertices are
    written as the analytic checker reads them back (world frame)."""
    with open(path, "w") as fh:
        n = len(table["kind"])
        fh.write("%d\n" % n)
        for i in range(n):
            fh.write("----\n")
            fh.write("%.17g,%.17g\n" % (table["origin"][i][0], table["origin"][i][1]))
            fh.write("0\n")
            v0, v1 = table["vert_off"][i], table["vert_off"][i + 1]
            fh.write("%d\n" % (v1 - v0))
            for v in range(v0, v1):
                fh.write("%.17g,%.17g\n" % (table["verts"][v][0], table["verts"][v][1]))


def timed_obstacles(m=32, seed=0xD2270006, lo=5.0, hi=45.0):
    """m time-parametrised obstacles (kinds 6 / 7, obstacle.cpp:805-875): a disc of radius_ whose
    centre moves along path rows (dx, dy, time) relative to origin_."""
    rng = np.random.default_rng(seed)
    origin = rng.uniform(lo, hi, (m, 2))
    radius = rng.uniform(0.5, 2.0, m)
    rows = rng.integers(2, 9, m)
    path_off = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(rows, out=path_off[1:])
    path = np.empty((int(path_off[-1]), 3), dtype=np.float64)
    for i in range(m):
        t = np.cumsum(rng.uniform(0.5, 4.0, rows[i])) + rng.uniform(0.0, 5.0)
        xy = np.cumsum(rng.uniform(-3.0, 3.0, (rows[i], 2)), axis=0)
        path[path_off[i]:path_off[i + 1], :2] = xy
        path[path_off[i]:path_off[i + 1], 2] = t
    kind = np.where(rng.random(m) < 0.5, 6, 7).astype(np.int32)
    return dict(kind=kind, used=np.ones(m, dtype=np.uint8), life_span=np.full(m, 1000000000000.0),
                origin=origin, radius=radius, path_off=path_off, path=path)


def timed_segments(n, seed, lo=0.0, hi=50.0, t_max=30.0):
    """n robot moves (x, y, time) -> (x, y, time), a few units long, either time order"""
    rng = np.random.default_rng(seed)
    s = np.empty((n, 3))
    s[:, 0] = rng.uniform(lo, hi, n)
    s[:, 1] = rng.uniform(lo, hi, n)
    s[:, 2] = rng.uniform(0.0, t_max, n)
    e = s + np.stack([rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.uniform(-6, 6, n)], 1)
    return s, e
