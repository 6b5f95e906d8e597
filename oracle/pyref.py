"""Driver for oracle/_ref/drrt_ref -- the REFERENCE'S OWN sources compiled by
`make -C oracle ref` (TEST INFRASTRUCTURE; see ref_driver.cpp).  Used only to
validate the plain-C oracle restatement and to generate golden fixtures."""
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "drrt_ref")
# the same driver linked against the product's drop-in shim (make -C oracle shim)
SHIM_BIN = os.path.join(_HERE, "_ref", "drrt_shim")


def available():
    return os.path.exists(BIN)


def _run(cmd, payload):
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as fh:
            fh.write(payload)
        subprocess.run([BIN, cmd, fin, fout], check=True, stdout=subprocess.DEVNULL)
        with open(fout, "rb") as fh:
            return fh.read()


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).tobytes()


def _tree_payload(pos, q, rad, metric, wrap):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 3)
    rad = np.ascontiguousarray(np.broadcast_to(np.asarray(rad, dtype=np.float64), (len(q),)))
    head = struct.pack("<iiqq", metric, 1 if wrap else 0, len(pos), len(q))
    return head + _f64(pos) + _f64(q) + _f64(rad), len(pos), len(q)


def structure(pos, metric=0, wrap=True):
    payload, n, _ = _tree_payload(pos, np.zeros((0, 3)), 0.0, metric, wrap)
    out = np.frombuffer(_run("structure", payload), dtype=np.int32).reshape(4, n)
    return dict(parent=out[0], left=out[1], right=out[2], split=out[3])


def range_batch(pos, q, rad, metric=0, wrap=True):
    """CSR sorted by id, as RefTree.range_batch"""
    payload, n, nq = _tree_payload(pos, q, rad, metric, wrap)
    raw = _run("range", payload)
    total = struct.unpack_from("<q", raw, 0)[0]
    counts = np.frombuffer(raw, dtype=np.int32, count=nq, offset=8)
    ids = np.frombuffer(raw, dtype=np.int32, count=total, offset=8 + 4 * nq).copy()
    dist = np.frombuffer(raw, dtype=np.float64, count=total, offset=8 + 4 * nq + 4 * total).copy()
    off = np.zeros(nq + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    for i in range(nq):
        sl = slice(off[i], off[i + 1])
        order = np.argsort(ids[sl], kind="stable")
        ids[sl] = ids[sl][order]
        dist[sl] = dist[sl][order]
    return off, ids, dist


def recreate(pos, q, rad, metric=0, wrap=True):
    """Tree A over the first half of pos is queried and destroyed, tree B over all of pos is
    built and queried (ThetaStar's create / destroy / create pattern).  -> two dicts
    (root id, tree_size_, offsets, ids with each row sorted by id)."""
    payload, n, nq = _tree_payload(pos, q, rad, metric, wrap)
    raw = _run("recreate", payload)
    o = 0
    out = []
    for _ in range(2):
        total, root, size = struct.unpack_from("<qii", raw, o)
        o += 16
        counts = np.frombuffer(raw, dtype=np.int32, count=nq, offset=o)
        o += 4 * nq
        ids = np.frombuffer(raw, dtype=np.int32, count=total, offset=o).copy()
        o += 4 * total
        off = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        for i in range(nq):
            ids[off[i]:off[i + 1]].sort()
        out.append(dict(root=root, size=size, off=off, ids=ids))
    return out


def nearest_batch(pos, q, metric=0, wrap=True):
    payload, n, nq = _tree_payload(pos, q, 0.0, metric, wrap)
    raw = _run("nearest", payload)
    return (np.frombuffer(raw, dtype=np.int32, count=nq).copy(),
            np.frombuffer(raw, dtype=np.float64, count=nq, offset=4 * nq).copy())


def ghosts(q, best_dist, metric=0):
    payload, _, nq = _tree_payload(np.zeros((1, 3)), q, best_dist, metric, True)
    raw = _run("ghosts", payload)
    counts = np.frombuffer(raw, dtype=np.int32, count=nq)
    g = np.frombuffer(raw, dtype=np.float64, offset=4 * nq).reshape(-1, 3)
    return counts.copy(), g.copy()


def _obs_payload(o):
    kind = np.ascontiguousarray(o["kind"], dtype=np.int32)
    n = len(kind)
    return (struct.pack("<i", n) + kind.tobytes()
            + np.ascontiguousarray(o["used"], dtype=np.uint8).tobytes()
            + _f64(o["life_span"]) + _f64(o["origin"]) + _f64(o["radius"])
            + np.ascontiguousarray(o["vert_off"], dtype=np.int32).tobytes() + _f64(o["verts"]))


EMPTY_OBS = dict(kind=[], used=[], life_span=[], origin=np.zeros((0, 2)), radius=[], vert_off=[0],
                 verts=np.zeros((0, 2)))


def edges(start, end, obstacles=None, r_min=1.0, robot_radius=0.5, metric=0, want_traj=False,
          cmd=None):
    s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
    e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
    n = len(s)
    payload = (struct.pack("<idd", metric, r_min, robot_radius)
               + _obs_payload(obstacles or EMPTY_OBS) + struct.pack("<q", n) + _f64(s) + _f64(e))
    raw = _run(cmd or ("traj" if want_traj else "edgecheck"), payload)
    types = [raw[4 * i:4 * i + 3].decode() for i in range(n)]
    o = 4 * n
    dist = np.frombuffer(raw, dtype=np.float64, count=n, offset=o).copy()
    o += 8 * n
    npts = np.frombuffer(raw, dtype=np.int32, count=n, offset=o).copy()
    o += 4 * n
    verdict = np.frombuffer(raw, dtype=np.uint8, count=n, offset=o).copy()
    o += n
    out = dict(type=types, dist=dist, npts=npts, verdict=verdict)
    if want_traj:
        xy = np.frombuffer(raw, dtype=np.float64, offset=o).reshape(-1, 2)
        offs = np.concatenate([[0], np.cumsum(npts)])
        out["xy"] = [xy[offs[i]:offs[i + 1]].copy() for i in range(n)]
    return out


def points(pts, obstacles, robot_radius=0.5, collision_distance=0.1, metric=0):
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    payload = (struct.pack("<idd", metric, robot_radius, collision_distance)
               + _obs_payload(obstacles) + struct.pack("<q", len(p)) + _f64(p))
    return np.frombuffer(_run("points", payload), dtype=np.uint8, count=len(p)).copy()


def geom(pa, pb, qa, qb):
    v = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1, 2) for x in (pa, pb, qa, qb)],
                       axis=1)
    n = len(v)
    raw = _run("geom", struct.pack("<q", n) + _f64(v))
    return (np.frombuffer(raw, dtype=np.float64, count=n).copy(),
            np.frombuffer(raw, dtype=np.float64, count=n, offset=8 * n).copy())


def obstacle_file_edges(path, start, end, r_min=1.0, robot_radius=0.5, metric=0):
    """reads `path` with the reference's Obstacle::ReadObstaclesFromFile, then checks the edges"""
    s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
    e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
    pb = os.path.abspath(path).encode()
    payload = (struct.pack("<iddi", metric, r_min, robot_radius, len(pb)) + pb
               + struct.pack("<q", len(s)) + _f64(s) + _f64(e))
    raw = _run("obsfile", payload)
    n_obs = struct.unpack_from("<i", raw, 0)[0]
    verdict = np.frombuffer(raw, dtype=np.uint8, count=len(s), offset=4).copy()
    dist = np.frombuffer(raw, dtype=np.float64, count=len(s), offset=4 + len(s)).copy()
    return n_obs, verdict, dist


# ---- timing commands (bench.py: the reference itself as the CPU baseline) ------------------
def range_time_payload(pos, q, rad, warm, timed, metric=0, wrap=True):
    payload, _, _ = _tree_payload(pos, q, rad, metric, wrap)
    return payload + struct.pack("<ii", int(warm), int(timed))


def edge_time_payload(start, end, obstacles, r_min=1.0, robot_radius=0.5, metric=0):
    s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
    e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
    return (struct.pack("<idd", metric, r_min, robot_radius) + _obs_payload(obstacles or EMPTY_OBS)
            + struct.pack("<q", len(s)) + _f64(s) + _f64(e))


def parse_time(raw):
    """-> (hits, seconds) written by `range_time` (hits of the first timed pass) / `edgecheck_time`"""
    return struct.unpack_from("<qd", raw, 0)


def range_time(pos, q, rad, warm=0, timed=1, metric=0, wrap=True):
    return parse_time(_run("range_time", range_time_payload(pos, q, rad, warm, timed, metric, wrap)))


def edge_time(start, end, obstacles, r_min=1.0, robot_radius=0.5, metric=0):
    return parse_time(_run("edgecheck_time", edge_time_payload(start, end, obstacles, r_min, robot_radius, metric)))


def loop_time(obstacles, iters, mode=0, seed=0xD2270011, delta=5.0, gamma=100.0, r_min=1.0,
              robot_radius=0.5, metric=0, binary=None):
    """`loop_time`: the planner iteration of BASELINE config 1 over the reference's own objects
    (ref_driver.cpp cmd_loop).  binary: BIN (all-CPU reference) or SHIM_BIN (tree calls and
    checks on the GPU; mode 0 = one ExplicitEdgeCheck call per edge, 1 = one call per iteration).
    -> dict(nodes, hits, edges, collisions, blocked, key_sum, len_sum, seconds)"""
    global BIN
    payload = (struct.pack("<idd", metric, r_min, robot_radius) + _obs_payload(obstacles or EMPTY_OBS)
               + struct.pack("<iiqdd", int(iters), int(mode), int(seed), delta, gamma))
    old = BIN
    if binary:
        BIN = binary
    try:
        raw = _run("loop_time", payload)
    finally:
        BIN = old
    v = struct.unpack_from("<qqqqqqdd", raw, 0)
    return dict(zip(("nodes", "hits", "edges", "collisions", "blocked", "key_sum", "len_sum", "seconds"), v))


def timed(obstacles, start, end, radius=0.5):
    """ExplicitEdgeCheck2D with kind 6 / 7 obstacles (obstacle.cpp:805-875): verdict per (x, y, time) segment,
    OR over the obstacles.  obstacles: dict(kind, used, life_span, origin, radius, path_off, path)"""
    o = obstacles
    kind = np.ascontiguousarray(o["kind"], dtype=np.int32)
    s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
    e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
    payload = (struct.pack("<di", radius, len(kind)) + kind.tobytes()
               + np.ascontiguousarray(o["used"], dtype=np.uint8).tobytes() + _f64(o["life_span"]) + _f64(o["origin"])
               + _f64(o["radius"]) + np.ascontiguousarray(o["path_off"], dtype=np.int32).tobytes() + _f64(o["path"])
               + struct.pack("<q", len(s)) + _f64(s) + _f64(e))
    return np.frombuffer(_run("timed", payload), dtype=np.uint8, count=len(s)).copy()
