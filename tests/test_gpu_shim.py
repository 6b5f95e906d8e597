"""Drop-in check: the REFERENCE'S OWN compiled objects (JList, Edge, Obstacle,
the driver that calls KDTree's public methods) linked against the product's
shim (drrt_b200/shim/kdtree_gpu.cpp, obstacle_gpu.cpp) instead of the
reference's kdtree.cpp -- `make -C oracle shim` -> oracle/_ref/drrt_shim.  The
reference-side code then runs unchanged and every tree call lands in the CUDA
library; results must equal the all-CPU reference binary's golden outputs."""
import json
import os

import numpy as np
import pytest

from test_oracle_vs_ref import gen, unhex

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(drrt):
    from oracle import pyref
    if not os.path.exists(pyref.SHIM_BIN):
        pytest.skip("oracle/_ref/drrt_shim not built (needs /root/reference at build time)")
    old = pyref.BIN
    pyref.BIN = pyref.SHIM_BIN
    yield pyref
    pyref.BIN = old


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "ref_v1.json")) as fh:
        return json.load(fh)


def test_reference_objects_over_gpu_tree(shim, gold):
    I = gen.inputs()
    st = shim.structure(I["pos"])              # kd_parent_/kd_child_*/kd_split_ mirrored from the device
    for k in ("parent", "left", "right", "split"):
        assert st[k].tolist() == gold["structure"][k], k
    off, ids, dist = shim.range_batch(I["pos"], I["q"], I["rad"])   # KDFindWithinRange -> JList -> pops
    g = gold["range_wrap_r2s1"]
    assert off.tolist() == g["off"] and ids.tolist() == g["ids"]
    assert np.array_equal(dist, unhex(g["dist"]))
    nid, nd = shim.nearest_batch(I["pos"], I["q"])
    assert nid.tolist() == gold["nearest_wrap_r2s1"]["ids"]
    off, ids, dist = shim.range_batch(I["grid"], I["grid"][::5], 2.0, metric=1)   # Theta*'s DistFunc tree
    g = gold["range_lattice"]
    assert off.tolist() == g["off"] and ids.tolist() == g["ids"]


def test_tree_destroyed_and_recreated(shim, po):
    """ThetaStar's pattern (thetastar.cpp:50-51 re-entered from obstacle.cpp:526): a local KDTree
    dies and the next make_shared<KDTree> usually gets its address back.  The shim keys its state
    by that address; the second tree must start from an empty device store (root = its own first
    node, ids from 0) and answer from its own nodes only."""
    I = gen.inputs()
    a, b = shim.recreate(I["pos"], I["q"], I["rad"])
    n = len(I["pos"])
    for got, m in ((a, n // 2), (b, n)):
        ref = po.RefTree()
        ref.insert(I["pos"][:m])
        roff, rids, _ = ref.range_batch(I["q"], I["rad"])
        assert got["root"] == 0 and got["size"] == m
        assert np.array_equal(got["off"], roff) and np.array_equal(got["ids"], rids)


def test_reference_edges_over_gpu_checker(shim, po, gold):
    I = gen.inputs()
    r = shim.edges(I["s"], I["e"], I["o"], cmd="edgecheck_gpu")
    want = np.array(gold["edges"]["verdict"], dtype=np.uint8)
    diff = np.nonzero(r["verdict"] != want)[0]
    if len(diff):
        _, _, clr = po.RefObstacles(**I["o"]).edge_check_batch(I["s"], I["e"], want_clearance=True)
        assert (clr[diff] < 1e-6).all()
    assert np.allclose(r["dist"], unhex(gold["edges"]["dist"]), rtol=1e-9)


def test_obstacle_file_to_device_table(shim, po, tmp_path):
    """the reference's own Obstacle::ReadObstaclesFromFile (obstacle.cpp:25-155)
    fills ConfigSpace::obstacles_; DrrtGpuUploadObstacles flattens that list into the device table;
    ExplicitEdgeCheckGpu then answers as the analytic checker does on the same list."""
    from drrt_b200 import synth
    o = synth.obstacles(40, 0xE1)
    path = str(tmp_path / "obstacles.txt")
    synth.write_obstacle_file(path, o)
    s, e = synth.random_edges(1500, 0xE2, max_len=2.5)
    n_obs, v, d = shim.obstacle_file_edges(path, s, e)
    rv, rl, clr = po.RefObstacles(**o).edge_check_batch(s, e, want_clearance=True)
    assert n_obs == 40
    bad = np.nonzero(v != rv)[0]
    assert (clr[bad] < 1e-6).all() and len(bad) <= 2
    assert np.allclose(d, rl, rtol=1e-9)
    assert 0.05 < v.mean() < 0.9


def test_planner_loop_reference_vs_shim(shim):
    """BASELINE config 1 through the reference's own objects: the same driver loop (sample ->
    KDFindNearest -> Saturate -> ExplicitNodeCheck -> KDFindWithinRange -> 2|N| edges -> KDInsert,
    ref_driver.cpp cmd_loop) once over the reference's kdtree.cpp + analytic checker
    (oracle/_ref/drrt_ref, all CPU) and once over the shim + libdrrt_b200.so: the same nodes get
    inserted, every neighbour list has the same keys, the same edges collide."""
    from drrt_b200 import synth
    o = synth.obstacles(24, 0xD2270012, lo=8.0, hi=42.0)
    iters = 1500
    cpu = shim.loop_time(o, iters, binary=shim.BIN.replace("drrt_shim", "drrt_ref"))
    for mode in (0, 1):
        gpu = shim.loop_time(o, iters, mode=mode, binary=shim.SHIM_BIN)
        for k in ("nodes", "hits", "edges", "blocked", "key_sum"):
            assert gpu[k] == cpu[k], (mode, k, gpu[k], cpu[k])
        assert abs(gpu["collisions"] - cpu["collisions"]) <= 2          # edges within 1e-6 of a threshold
        assert abs(gpu["len_sum"] - cpu["len_sum"]) <= 1e-9 * cpu["len_sum"]
    assert cpu["nodes"] > 1000 and cpu["edges"] > 20_000 and 0 < cpu["collisions"] < cpu["edges"]
