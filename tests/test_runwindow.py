"""Host-side model of the run lookup in scan_columns_cta (drrt_b200/csrc/gridscan.cuh).

The CTA-wide range scan numbers the candidates of a query 0..total-1 over a list of non-empty
record runs (runs[j] = (exclusive end of run j in candidate numbers, record offset)); a warp
walks a contiguous share [c0, c1) in trips of 32 candidates.  The kernel finds the run of
candidate i0+lane from a window of 32 runs held one per lane: one warp-wide OR marks the
candidate numbers at which runs of the window end, lane L's run is the population count of the
marks at or below L, and the next window starts popcount(marks) (+1 when a run ends exactly at
i0+32) runs further.  This test replays that arithmetic in numpy, lane by lane, against a
plain searchsorted over the run list -- including runs of length 1, windows that reach the
sentinel and shares that end inside a trip.
"""
import numpy as np
import pytest

SENTINEL = 0x7FFFFFFF


def lookup_trips(ends, offs, c0, c1):
    """record index of every candidate of [c0, c1), the way the kernel computes it"""
    nent = len(ends)
    ends_s = np.concatenate([ends, [SENTINEL]]).astype(np.int64)
    offs_s = np.concatenate([offs, [0]]).astype(np.int64)
    jl = int(np.searchsorted(ends, c0, side="right"))  # smallest j with ends[j] > c0
    out = np.full(c1 - c0, -1, dtype=np.int64)
    lanes = np.arange(32)
    for i0 in range(c0, c1, 32):
        idx = np.minimum(jl + lanes, nent)
        ex, ey = ends_s[idx], offs_s[idx]
        rel = (ex - i0).astype(np.uint64)  # the kernel's unsigned subtraction
        assert (rel >= 1).all(), "run jl must hold candidate i0"
        marks = 0
        for r in rel:
            if r < 32:
                assert not (marks >> int(r)) & 1, "two runs of a window end at the same candidate"
                marks |= 1 << int(r)
        for lane in lanes:
            i = i0 + lane
            mask = ((2 << lane) - 1) & 0xFFFFFFFF
            k = bin(marks & mask).count("1")
            assert k <= lane
            if i < c1:
                out[i - c0] = i + ey[k]
        jl += bin(marks).count("1") + (1 if (rel == 32).any() else 0)
    return out


def naive(ends, offs, c0, c1):
    i = np.arange(c0, c1)
    j = np.searchsorted(ends, i, side="right")
    return i + offs[j]


def make_runs(rng, nrun, max_len):
    lens = rng.integers(1, max_len + 1, nrun)
    ends = np.cumsum(lens)
    starts = rng.integers(0, 1_000_000, nrun)  # first record of each run
    offs = starts - (ends - lens)              # record = candidate + offs
    return ends.astype(np.int64), offs.astype(np.int64)


@pytest.mark.parametrize("max_len", [1, 2, 19, 40, 500])
def test_window_lookup_matches_searchsorted(max_len):
    rng = np.random.default_rng(0xD227_1000 + max_len)
    for _ in range(20):
        nrun = int(rng.integers(1, 600))
        ends, offs = make_runs(rng, nrun, max_len)
        total = int(ends[-1])
        nwarps = 16
        per = ((total + nwarps - 1) // nwarps + 31) & ~31
        for w in range(nwarps):
            c0, c1 = w * per, min(total, (w + 1) * per)
            if c0 >= c1:
                continue
            assert np.array_equal(lookup_trips(ends, offs, c0, c1), naive(ends, offs, c0, c1))


def test_window_lookup_exact_multiples():
    # every run exactly 32 long: each trip's runs end at i0 + 32 (the vote branch)
    ends = np.arange(1, 41, dtype=np.int64) * 32
    offs = np.arange(40, dtype=np.int64) * 1000
    assert np.array_equal(lookup_trips(ends, offs, 0, 1280), naive(ends, offs, 0, 1280))
    # one long run, then 64 runs of length 1
    ends = np.concatenate([[100], 100 + np.arange(1, 65)]).astype(np.int64)
    offs = np.arange(65, dtype=np.int64) * 7
    assert np.array_equal(lookup_trips(ends, offs, 64, 164), naive(ends, offs, 64, 164))
