"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's AccumulatePass and ErrorMeasurePass arithmetic.
Only tests/ may import this.  Pinned by hand-computed known-answer cases in tests/test_post_oracle.py (the reference holds
no fixtures for these passes).

  accumulate_*   Source/RenderPasses/AccumulatePass/Accumulate.cs.slang:57-122 (+ the clear at frame 0, AccumulatePass.cpp)
  error_measure  Source/RenderPasses/ErrorMeasurePass/ErrorMeasurer.cs.slang:41-61, ErrorMeasurePass.cpp:241-243
"""
import numpy as np

F = np.float32


class Accumulator:
    def __init__(self, mode="Double"):
        self.mode = mode
        self.count = 0
        self.sum = None
        self.corr = None

    def add(self, cur):
        cur = np.asarray(cur, dtype=F)
        if self.count == 0:
            self.sum = np.zeros(cur.shape, dtype=np.float64 if self.mode == "Double" else F)
            self.corr = np.zeros(cur.shape, dtype=F)
        n = self.count + 1
        if self.mode == "Single":                      # :57-70
            self.sum = (self.sum + cur).astype(F)
            out = (self.sum / F(n)).astype(F)
        elif self.mode == "SingleCompensated":         # :74-93
            y = (cur - self.corr).astype(F)
            nxt = (self.sum + y).astype(F)
            out = (nxt / F(n)).astype(F)
            self.corr = ((nxt - self.sum).astype(F) - y).astype(F)
            self.sum = nxt
        else:                                          # :97-122
            self.sum = self.sum + cur.astype(np.float64)
            out = (self.sum / np.float64(n)).astype(F)
        self.count += 1
        return out


def error_measure(source, reference, world_position=None, ignore_background=True, squared=True, average=False):
    s = np.asarray(source, dtype=F)[..., :3]
    r = np.asarray(reference, dtype=F)[..., :3]
    d = np.abs((s - r).astype(F))
    if ignore_background and world_position is not None:
        d = np.where((np.asarray(world_position)[..., 3] != 0)[..., None], d, F(0))
    if squared:
        d = (d * d).astype(F)
    if average:
        a = (((d[..., 0] + d[..., 1]).astype(F) + d[..., 2]).astype(F) / F(3)).astype(F)
        d = np.stack([a, a, a], axis=-1)
    total = d.reshape(-1, 3).astype(np.float64).sum(axis=0)
    npx = F(s.shape[0] * s.shape[1])
    err = (total.astype(F) / npx).astype(F)
    avg = (((err[0] + err[1]).astype(F) + err[2]).astype(F) / F(3)).astype(F)
    return d, err, avg
