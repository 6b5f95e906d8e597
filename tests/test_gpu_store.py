"""KDInsert parity: the device builds the same insertion-ordered KD tree
(parent / children / split) as KDTree::KDInsert, src/kdtree.cpp:87-136."""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu


def _same_structure(t, r):
    a, b = t.structure(), r.structure()
    for k in ("parent", "left", "right", "split"):
        assert np.array_equal(a[k], b[k]), k


def test_insert_one_batch(drrt, po):
    pos = synth.box_points(50_000, 5)
    t = drrt.KDTree(); r = po.RefTree()
    assert t.KDInsert(pos) == 0
    r.insert(pos)
    assert t.tree_size_ == 50_000
    _same_structure(t, r)
    assert np.array_equal(t.positions(), pos)


def test_insert_batch_split_invariance(drrt, po):
    pos = synth.box_points(20_000, 6)
    pos[5000:5100] = pos[4000:4100]      # exact duplicates go right (>=)
    pos[7000:7050, 0] = pos[0, 0]        # ties on the root's split coordinate
    r = po.RefTree(); r.insert(pos)
    t = drrt.KDTree()
    done = 0
    for chunk in (1, 1, 1, 5, 1000, 1024, 1025, 3000, 1, 13942):
        assert t.KDInsert(pos[done:done + chunk]) == done
        done += chunk
    assert done == len(pos)
    _same_structure(t, r)


def test_insert_sorted_input_deep_tree(drrt, po):
    # monotone x,y,theta: the tree degenerates to a path of depth n
    n = 600
    pos = np.stack([np.linspace(0, 50, n), np.linspace(0, 50, n), np.linspace(0, 6, n)], 1)
    t = drrt.KDTree(); r = po.RefTree()
    t.KDInsert(pos); r.insert(pos)
    _same_structure(t, r)


def test_insert_rejects_nonfinite(drrt):
    t = drrt.KDTree()
    with pytest.raises(drrt.DrrtError):
        t.KDInsert(np.array([[0.0, np.nan, 0.0]]))


def test_create_rejects_unsupported(drrt):
    with pytest.raises(drrt.DrrtError):
        drrt.KDTree(dims=2)
    with pytest.raises(drrt.DrrtError):
        drrt.KDTree(wraps=(0,), wrap_points=(50.0,))
    with pytest.raises(drrt.DrrtError):
        drrt.KDTree(wraps=(2,), wrap_points=(6.0,))   # R2S1 metric wraps at 2*pi' only


def test_one_handle_many_host_threads(drrt, po):
    # the reference enters the tree from 5 planner threads + robot + obstacle threads under
    # tree_mutex_ (mainloop.cpp:99-113); calls on one handle are serialised inside the library
    import threading
    pos = synth.box_points(40_000, 7)
    t = drrt.KDTree(); t.KDInsert(pos)
    r = po.RefTree(); r.insert(pos)
    q = synth.box_points(4 * 300, 8).reshape(4, 300, 3)
    want = [r.range_batch(q[k], 1.5) for k in range(4)]
    got = [None] * 4
    errs = []

    def work(k):
        try:
            for _ in range(5):
                got[k] = t.KDFindWithinRange(1.5, q[k])
                nid, _ = t.KDFindNearest(q[k][:50])
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs
    for k in range(4):
        for a, b in zip(got[k], want[k]):
            assert np.array_equal(a, b)
