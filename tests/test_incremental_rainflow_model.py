"""The incremental rainflow / SEI scheme of the post kernel (fleetstep.cu rf_vehicle), restated in
Python and checked against the reference semantics: rainflow.extract_cycles over the WHOLE history at every
evaluation (oracle/fleet_oracle.c rainflow_cycles, itself pinned by the ASTM KAT and the reference goldens) followed by
the positional slice [rainflow_length-1 : len-1] of rainflow_sei_degradation.py:143-146.

What must hold for any series, any split of it into consumed batches and any evaluation times:
  * cycles committed incrementally == the stable prefix of the reference's cycle list,
  * m (list length) and the slice's (range, mean, count) multiset == the reference's,
  * the sum of all cycle means == the reference's.
This is the argument the CUDA kernel relies on; the kernel itself is checked on the GPU against the oracle.
"""
import zlib

import numpy as np
import pytest

from oracle.oracle import rainflow_cycles


class IncrementalRainflow:
    """Per-vehicle committed state: stack of open reversal points, committed cycle count, running sums."""

    def __init__(self, x0):
        self.stack = [x0]
        self.c = 0
        self.msum = 0.0
        self.pend = []          # committed cycles at list positions >= rf_len-1 (the kernel keeps their stress sum)
        self.x_cur = x0
        self.n = 1              # samples seen

    def consume(self, samples, rf_len):
        """reversal detection continues from the last consumed sample; direction = sign(x_cur - top of stack)"""
        dsg = self.x_cur - self.stack[-1]
        for x_next in samples:
            self.n += 1
            if x_next != self.x_cur:
                d = x_next - self.x_cur
                if (dsg < 0 < d) or (dsg > 0 > d):
                    self._push(self.x_cur, rf_len)
                dsg = d
                self.x_cur = x_next

    def _commit(self, xa, xb, cnt, rf_len):
        self.msum += 0.5 * (xa + xb)
        if self.c >= rf_len - 1:
            self.pend.append((abs(xa - xb), 0.5 * (xa + xb), cnt))
        self.c += 1

    def _push(self, v, rf_len):
        st = self.stack
        st.append(v)
        while len(st) >= 3:
            x1, x2, x3 = st[-3], st[-2], st[-1]
            if abs(x3 - x2) < abs(x2 - x1):
                break
            if len(st) == 3:
                self._commit(x1, x2, 0.5, rf_len)
                st.pop(0)
            else:
                self._commit(x1, x2, 1.0, rf_len)
                del st[-3:-1]

    def evaluate(self, rf_len):
        """-> (m, sum of all means, slice cycles, new rf_len); the stack is only read"""
        m, msum, sl = self.c, self.msum, list(self.pend)
        if self.n >= 3:
            st, x = self.stack, self.x_cur
            h, lo = len(st), 0
            prov = []
            while h - lo >= 2:
                x2, x1 = st[h - 1], st[h - 2]
                if abs(x - x2) < abs(x2 - x1):
                    break
                if h - lo == 2:
                    prov.append((x1, x2, 0.5)); lo += 1
                else:
                    prov.append((x1, x2, 1.0)); h -= 2
            for k in range(lo, h - 1):
                prov.append((st[k], st[k + 1], 0.5))
            prov.append((st[h - 1], x, 0.5))
            for j, (xa, xb, cnt) in enumerate(prov):
                msum += 0.5 * (xa + xb)
                if j < len(prov) - 1 and m >= rf_len - 1:
                    sl.append((abs(xa - xb), 0.5 * (xa + xb), cnt))
                m += 1
        if m > rf_len:
            self.pend = []
            return m, msum, sl, m
        assert not self.pend
        return m, msum, None, rf_len


def _series(rng, n, kind):
    if kind == "walk":
        x = np.clip(0.5 + np.cumsum(rng.normal(0, 0.05, n)), 0, 1)
    elif kind == "plateaus":
        x = np.clip(0.5 + np.cumsum(rng.normal(0, 0.05, n) * (rng.random(n) < 0.4)), 0, 1)
    elif kind == "damped":          # strictly shrinking swings: the stack grows with every reversal
        k = np.arange(n)
        x = 0.5 + 0.45 * (-1.0) ** k * 0.97 ** k
    elif kind == "grid":            # few distinct values: many exact ties (X == Y closes a cycle)
        x = rng.integers(0, 5, n) / 4.0
    else:
        x = np.full(n, 0.37)
    return x.astype(np.float64)


@pytest.mark.parametrize("kind", ["walk", "plateaus", "damped", "grid", "flat"])
def test_incremental_equals_full_rescan(kind):
    rng = np.random.default_rng(zlib.crc32(kind.encode()) % 1000)
    for trial in range(60):
        n = int(rng.integers(2, 260))
        x = _series(rng, n, kind)
        rf_len = int(rng.integers(1, 12))             # carried over from earlier episodes
        inc = IncrementalRainflow(x[0])
        pos = 1
        while pos < n:
            step = int(rng.integers(1, 40))
            nxt = min(n, pos + step)
            inc.consume(x[pos:nxt], rf_len)
            pos = nxt
            if rng.random() < 0.5 or pos == n:        # evaluation with the history x[:pos]
                ref = rainflow_cycles(x[:pos])
                m, msum, sl, new_len = inc.evaluate(rf_len)
                assert m == len(ref)
                np.testing.assert_allclose(msum, sum(c[1] for c in ref), rtol=1e-13, atol=1e-13)
                if len(ref) > rf_len:
                    want = sorted((c[0], c[1], c[2]) for c in ref[rf_len - 1:len(ref) - 1])
                    assert sl is not None and sorted(sl) == want
                    rf_len = len(ref)
                else:
                    assert sl is None
                assert new_len == rf_len
            # the committed cycles are a stable prefix of the reference's list
            full = rainflow_cycles(x[:pos])
            assert inc.c <= max(len(full) - 1, 0) or pos < 3
