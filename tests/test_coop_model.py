"""The lane-level Python model of the cooperative kernels (tests/coop_model.py): plan coverage,
carry resolution with generate/propagate lookahead and the three-phase Montgomery product against
plain integer arithmetic.  The CUDA code in csrc/dkg_coop.cuh is a transcription of this model."""
from __future__ import annotations

import random

import coop_model as cm


def test_three_phase_montgomery_product_and_plans():
    cm.test_selfcheck()


def test_add_sub_lookahead_against_integers():
    rng = random.Random(2)
    K = 6
    W = 1 << (32 * K)
    for _ in range(200):
        nb = rng.choice([1, 3, 11, 16])
        top = W**nb
        # values with long runs of all-ones blocks exercise the propagate path
        x = rng.choice([rng.randrange(top), top - 1, top - rng.randrange(1, 5), (top - 1) ^ (W - 1)])
        y = rng.choice([rng.randrange(top), 1, W % top, top - 1])
        xs, ys = cm.blocks_of(x, nb, K), cm.blocks_of(y, nb, K)
        s, _ = cm.add(xs, ys, K)
        assert cm.value_of(s, K) == x + y
        d, borrow = cm.sub(xs, ys, K)
        assert cm.value_of(d, K) % (W**cm.LANES) == (x - y) % (W**cm.LANES)
        assert borrow == (1 if x < y else 0)
