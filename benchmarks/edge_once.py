#!/usr/bin/env python
"""One device step of BASELINE config 3 (10 M Dubins edges = the query<->hit pairs of the 2b range
batch, vs 256 polygons) and nothing else after the set-up: the command ncu wraps to count the FP64
work of the edge-check kernels (profiles/r02_edge_fp64_ops.json) and to list their launches.

    python benchmarks/edge_once.py [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import drrt_b200  # noqa: E402
from drrt_b200 import synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
nodes, q = synth.nodes(), synth.queries()
tree = drrt_b200.KDTree(device=0, capacity=len(nodes))
tree.KDInsert(nodes)
off, ids, _ = tree.KDFindWithinRange(synth.shrinking_radius(len(nodes), synth.GAMMA_SPARSE), q)
s_h, e_h = synth.edges_from_neighbours(nodes, q, off, ids, 10_000_000)
tree.SetObstacles(**synth.obstacles())
start, end = torch.from_numpy(s_h).to(dev), torch.from_numpy(e_h).to(dev)
verdict = torch.empty(len(s_h), dtype=torch.uint8, device=dev)
length = torch.empty(len(s_h), dtype=torch.float64, device=dev)
for _ in range(steps):
    tree.ExplicitEdgeCheck_dev(verdict, start=start, end=end, robot_radius=synth.ROBOT_RADIUS, r_min=synth.R_MIN,
                               length_out=length)
torch.cuda.synchronize()
print("edges", len(s_h), "collide", int(verdict.sum().item()))
