"""The synthetic-scene generators exist in the product (b2n/synthetic.py: benchmarks, GPU tests) and in the oracle
(oracle/nerf_oracle.py: CPU reference arm, golden scripts) because neither side may import the other; they must stay
bit-identical so that both arms of bench.py and the parity tests see the same rays."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "project-nerf_b200")]


def test_product_and_oracle_generators_are_identical():
    from b2n import synthetic as S
    from oracle import nerf_oracle as O
    for seed, n, with_time in ((0, 257, False), (3, 1000, False), (5, 500, True)):
        a = O.synthetic_rays(n, seed=seed, with_time=with_time)
        b = S.random_rays(n, seed=seed, with_time=with_time)
        assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(O.synthetic_poses(9, seed=2), S.hemisphere_poses(9, seed=2))
    for R, bound, radius in ((32, 1.5, 0.75), (48, 1.5, 1.0)):
        assert torch.equal(O.ball_occupancy(R, bound, radius), S.ball_occupancy(R, bound, radius))
