"""Restatement of `rainflow==3.2.0` (PyPI; pinned in the reference's requirements.txt, not vendored
under /root/reference and not installed here): ASTM E1049-85 three-point rainflow counting.

TEST INFRASTRUCTURE ONLY. Written from the published algorithm as summarised in SURVEY.md A.6;
pinned by the package README's ASTM example (tests/test_rainflow_kat.py).  Call sites in the
reference: fleetrl/utils/battery_degradation/rainflow_sei_degradation.py:1,132.
"""
from collections import deque


def reversals(series):
    """Yield (index, value) of the reversal points; first and last samples are always reversals,
    plateaus are skipped with an exact == test, and the index reported for a plateau is its last sample."""
    it = iter(series)
    x_last = next(it, None)
    x = next(it, None)
    if x_last is None or x is None:
        return
    d_last = x - x_last
    yield 0, x_last
    index = None
    x_next = None
    for index, x_next in enumerate(it, start=1):
        if x_next == x:
            continue
        d_next = x_next - x
        if d_last * d_next < 0:
            yield index, x
        x_last, x = x, x_next
        d_last = d_next
    if index is not None:
        yield index + 1, x_next


def extract_cycles(series):
    """Yield (range, mean, count, i_start, i_end); count is 0.5 (half cycle) or 1.0 (full cycle)."""
    points = deque()

    def fmt(p1, p2, count):
        i1, x1 = p1
        i2, x2 = p2
        return abs(x1 - x2), 0.5 * (x1 + x2), count, i1, i2

    for point in reversals(series):
        points.append(point)
        while len(points) >= 3:
            x1, x2, x3 = points[-3][1], points[-2][1], points[-1][1]
            X = abs(x3 - x2)
            Y = abs(x2 - x1)
            if X < Y:
                break
            elif len(points) == 3:
                yield fmt(points[0], points[1], 0.5)
                points.popleft()
            else:
                yield fmt(points[-3], points[-2], 1.0)
                last = points.pop()
                points.pop()
                points.pop()
                points.append(last)
    while len(points) > 1:
        yield fmt(points[0], points[1], 0.5)
        points.popleft()


def count_cycles(series, ndigits=None, nbins=None, binsize=None):
    from collections import defaultdict
    counts = defaultdict(float)
    for rng, _mean, count, _i, _j in extract_cycles(series):
        counts[rng if ndigits is None else round(rng, ndigits)] += count
    return sorted(counts.items())
