"""A/B timing of the config-3 edge checks (10 M Dubins edges vs 256 polygons) for one build of the
library (DRRT_B200_LIB selects it).  Prints ms per batch for Dubins and straight edges."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drrt_b200
from drrt_b200 import synth
from drrt_b200.kdtree import KDTree

dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
cache = "/tmp/edges_c3.npz"
if os.path.exists(cache):
    z = np.load(cache); s_h, e_h = z["s"], z["e"]
else:
    nodes = synth.nodes(); q = synth.queries()
    t0 = KDTree(device=0); t0.KDInsert(nodes)
    off, ids, _ = t0.KDFindWithinRange(synth.shrinking_radius(len(nodes), synth.GAMMA_SPARSE), q)
    s_h, e_h = synth.edges_from_neighbours(nodes, q, off, ids, 10_000_000)
    np.savez(cache, s=s_h, e=e_h); t0.close()
tree = KDTree(device=0)
tree.SetObstacles(**synth.obstacles())
start = torch.from_numpy(s_h).to(dev); end = torch.from_numpy(e_h).to(dev)
n = start.shape[0]
verdict = torch.empty(n, dtype=torch.uint8, device=dev); length = torch.empty(n, dtype=torch.float64, device=dev)
for kind, name in ((drrt_b200.EDGE_DUBINS, "dubins"), (drrt_b200.EDGE_STRAIGHT, "straight")):
    ts = []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tree.ExplicitEdgeCheck_dev(verdict, start=start, end=end, robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN, edge_kind=kind, length_out=length)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(os.environ.get("DRRT_B200_LIB", "default").split("/")[-2:], name, "ms %.3f (min of %s)" % (min(ts[2:]), ["%.2f" % t for t in ts]), "collide %.4f" % verdict.float().mean().item(), "checksum", int(verdict.sum().item()), flush=True)
