"""Stand-in for intervaltree.IntervalTree (from_tuples / merge_overlaps / iteration) - test infra.

Semantics restated from the intervaltree package's documented behaviour: intervals are half-open
[begin, end); merge_overlaps(strict=False) also merges intervals that merely touch (end == begin);
a null interval (begin >= end) raises ValueError.
"""
from collections import namedtuple

Interval = namedtuple("Interval", ["begin", "end", "data"])


class IntervalTree:
    def __init__(self, intervals=()):
        self._iv = set(intervals)

    @classmethod
    def from_tuples(cls, tups):
        ivs = []
        for t in tups:
            b, e = t[0], t[1]
            if b >= e:
                raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree: Interval(%r, %r)" % (b, e))
            ivs.append(Interval(b, e, t[2] if len(t) > 2 else None))
        return cls(ivs)

    def merge_overlaps(self, data_reducer=None, data_initializer=None, strict=True):
        if not self._iv:
            return
        srt = sorted(self._iv)
        merged = [srt[0]]
        for iv in srt[1:]:
            top = merged[-1]
            if iv.begin < top.end or (not strict and iv.begin == top.end):
                merged[-1] = Interval(top.begin, max(top.end, iv.end), None)
            else:
                merged.append(iv)
        self._iv = set(merged)

    def __iter__(self):
        return iter(self._iv)

    def __len__(self):
        return len(self._iv)
