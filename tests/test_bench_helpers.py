"""CPU checks of bench.py's bookkeeping: the roofline's `traffic` comes from the committed ncu
summary (never a literal), the FP64 work of config 3 from the committed ncu count, and stdout
carries exactly the one JSON line whatever libraries print to file descriptor 1."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_is_parsed_from_the_newest_profile():
    sys.path.insert(0, ROOT)
    import bench
    got = bench.ncu_traffic("range_cta_kernel")
    assert got is not None
    bytes_per_launch, source = got
    assert source.startswith("profiles/r02") and os.path.exists(os.path.join(ROOT, source))
    txt = open(os.path.join(ROOT, source)).read()
    assert "dram__bytes_read.sum" in txt and "dram__bytes_write.sum" in txt
    assert 2.5e9 < bytes_per_launch < 9e9          # rows out (2.8 GB) + what misses L2; below the algorithmic 8.4 GB
    assert bench.ncu_traffic("no_such_kernel") is None


def test_edge_fp64_ops_are_committed():
    sys.path.insert(0, ROOT)
    import bench
    ops = bench.edge_fp64_ops()
    assert ops and ops["fp64_flop_per_step"] == sum(k["fp64_flop"] for k in ops["kernels"].values())
    assert any("dubins_walk_kernel" in k for k in ops["kernels"])
    assert 500 < ops["fp64_flop_per_edge"] < 5000


def test_stdout_carries_one_json_line(tmp_path):
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('library chatter'); bench.emit({'metric': 'm', 'value': 1})" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1}
    assert "NCCL version banner" in out.stderr and "library chatter" in out.stderr
