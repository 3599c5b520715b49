"""Host-side replay of the 13-shuffle transposing butterfly of ``csrc/blend_bwd.cu`` (warp_transpose_reduce11 with
reduce11_slot / reduce11_writer): after the five exchange stages lane l must hold the warp total of value
``slot(l)``, and the writer lanes must cover the 11 values exactly once (they issue the single RED row update)."""
import numpy as np


def _slot(lane):
    s2 = 1 if lane & 2 else (2 if lane & 4 else 0)
    s1 = s2 + (3 if lane & 8 else 0)
    return (5 if s1 == 5 else s1 + 6) if lane & 16 else s1


def _writer(lane):
    if lane & 1:
        return False
    if (lane & 2) and (lane & 4):
        return False
    if (lane & 16) and (lane & 8) and not (lane & 2) and (lane & 4):
        return False
    return True


def _replay(v):  # v: (32 lanes, 11 values)
    v = v.astype(np.float64).copy()
    lanes = np.arange(32)

    def exchange(xor, pairs, both):
        new = v.copy()
        for lane in lanes:
            hi, partner = bool(lane & xor), lane ^ xor
            for keep_lo, keep_hi in pairs:  # the lane keeps one of the two slots and receives the partner's copy of it
                mine = keep_hi if hi else keep_lo
                new[lane, keep_lo] = v[lane, mine] + v[partner, mine]
            for slot in both:  # odd value of the stage: summed on both sides
                new[lane, slot] = v[lane, slot] + v[partner, slot]
        v[:] = new

    exchange(16, [(i, i + 6) for i in range(5)], [5])
    exchange(8, [(i, i + 3) for i in range(3)], [])
    exchange(4, [(0, 2)], [1])
    exchange(2, [(0, 1)], [])
    exchange(1, [], [0])
    return v[:, 0]


def test_butterfly_slots_and_writers():
    rng = np.random.default_rng(0)
    vals = rng.standard_normal((32, 11))
    out = _replay(vals)
    totals = vals.sum(axis=0)
    for lane in range(32):
        assert abs(out[lane] - totals[_slot(lane)]) < 1e-12
    writers = [_slot(lane) for lane in range(32) if _writer(lane)]
    assert sorted(writers) == list(range(11))
