"""Synthetic value fields for the MISE tests (shared by oracle/make_golden_mise.py and tests/test_mise.py).
Each takes (n,3) int64 lattice points and returns float64 values."""
import numpy as np


def known_answer(p, res):
    """The reference's own libmise/test.py field: MISE(1, 2, 0.) -> 3 rounds, 5^3 dense, sum 105.0."""
    return 2 * (p.sum(axis=-1) > 2).astype(np.float64) - 1


def sphere(p, res):
    return (0.3 - np.linalg.norm(p / res - 0.5, axis=-1)).astype(np.float32).astype(np.float64)


def rough(p, res):
    """Several blobs plus a ripple, quantised so that values exactly AT the threshold occur (they count as both
    positive and negative, mise.pyx:215-218) and refinement needs extra rounds."""
    c = np.array([[0.55, 0.72, 0.60], [0.54, 0.42, 0.65], [0.44, 0.89, 0.96], [0.38, 0.79, 0.53],
                  [0.57, 0.93, 0.07], [0.09, 0.02, 0.83]])
    x = p / res
    v = np.min(np.linalg.norm(x[:, None, :] - c[None], axis=-1), axis=1) - 0.17
    v = v + 0.03 * np.sin(37 * x[:, 0]) * np.cos(23 * x[:, 1] + 11 * x[:, 2])
    return (np.round(v * 8) / 8 * 0.25).astype(np.float64)


CASES = {  # name -> (resolution_0, depth, threshold, field)
    "known_answer_r1_d2": (1, 2, 0.0, known_answer),
    "sphere_r8_d2": (8, 2, 0.0, sphere),
    "sphere_r4_d4": (4, 4, 0.0, sphere),
    "rough_r8_d3": (8, 3, 0.0, rough),
    "rough_r16_d1_thr": (16, 1, 0.03125, rough),
}
