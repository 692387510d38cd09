"""Parity protocol shared by the CPU and GPU tests.

Codes are compared token by token.  A token whose codes equal the reference's on every layer is
*exact*.  Otherwise the first differing layer l* is looked up in the fp64 margin table computed along
the REFERENCE's code trajectory: margin < EPS means the two candidates were closer than any fp32
summation order can resolve (a *near-tie flip*; later layers of that token are excluded because the
recurrence has legitimately forked); margin >= EPS is a real disagreement and fails the test."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

EPS = 1e-5  # cos-sim units


@dataclass
class ParityReport:
    tokens: int
    exact: int
    near_tie: int
    failures: int
    worst_margin: float  # largest margin among accepted near-tie flips

    def __str__(self):
        return (f"{self.exact}/{self.tokens} tokens exact, {self.near_tie} near-tie flips "
                f"(max margin {self.worst_margin:.2e} < {EPS:g}), {self.failures} failures")


def compare_codes(test, ref, margins_ref) -> ParityReport:
    test = np.asarray(test).reshape(-1, np.asarray(test).shape[-1]).astype(np.int64)
    ref = np.asarray(ref).reshape(test.shape).astype(np.int64)
    m = np.asarray(margins_ref).reshape(test.shape)
    neq = test != ref
    bad_tok = np.flatnonzero(neq.any(axis=1))
    near, fail, worst = 0, 0, 0.0
    for t in bad_tok:
        l = int(np.argmax(neq[t]))
        if m[t, l] < EPS:
            near += 1
            worst = max(worst, float(m[t, l]))
        else:
            fail += 1
    return ParityReport(test.shape[0], test.shape[0] - len(bad_tok), near, fail, worst)


def exact_token_mask(test, ref):
    test = np.asarray(test)
    ref = np.asarray(ref).reshape(test.shape)
    return (test == ref).reshape(-1, test.shape[-1]).all(axis=1)
