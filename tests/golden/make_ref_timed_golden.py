"""Generates tests/golden/ref_timed_v1.json from oracle/_ref/drrt_ref: the reference's own
ExplicitEdgeCheck2D (src/obstacle.cpp:756-877) on time-parametrised obstacles (kinds 6 / 7,
:805-875, with FindIndexBeforeTime from src/drrt.cpp:178-189), compiled unmodified by
`make -C oracle ref`.  Inputs are seeded (drrt_b200.synth); only the verdicts are stored.

    make -C oracle ref && python tests/golden/make_ref_timed_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from drrt_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402


def inputs():
    o = synth.timed_obstacles(24, 0xC1)
    o["used"][3] = 0
    o["life_span"][4] = 0.0
    s, e = synth.timed_segments(4000, 0xC2)
    e[:50, 2] = s[:50, 2]                 # no extent in time: the parametric slopes divide by zero
    s[50:60] = e[50:60]                   # zero-length moves
    return dict(o=o, s=s, e=e, radii=(0.5, 1.25))


def main():
    assert pyref.available(), "build oracle/_ref first: make -C oracle ref"
    I = inputs()
    out = {"verdict": {str(r): pyref.timed(I["o"], I["s"], I["e"], r).tolist() for r in I["radii"]}}
    path = os.path.join(ROOT, "tests", "golden", "ref_timed_v1.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: int(np.sum(v)) for k, v in out["verdict"].items()}, "hits")


if __name__ == "__main__":
    main()
