"""Regenerates tests/golden/oracle_v1.json from the CPU oracle.

The reference has no golden vectors of its own; this fixture pins
the oracle restatement so that later edits to oracle/ cannot drift silently.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
from test_oracle import golden_payload  # noqa: E402

if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "oracle_v1.json")
    with open(out, "w") as fh:
        json.dump(golden_payload(po), fh)
    print("wrote", out, os.path.getsize(out), "bytes")
