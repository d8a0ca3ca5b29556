"""Synthetic value volumes for the marching-cubes tests (shared by oracle/make_golden_mcubes.py and tests/test_mcubes.py)."""
import numpy as np


def cases():
    rng = np.random.RandomState(3)
    x = np.linspace(-1, 1, 17)
    g = np.stack(np.meshgrid(x, x, x, indexing="ij"), -1)
    blob = 0.55 - np.linalg.norm(g * np.array([1.0, 1.3, 0.8]), axis=-1) + 0.12 * np.sin(7 * g[..., 0]) * np.cos(5 * g[..., 1])
    return {
        "blob_17": (blob, 0.0),                                                  # smooth closed surface
        "noise_9x7x6": (rng.randn(9, 7, 6), 0.1),                                # every cube configuration, anisotropic shape
        "quantised_8": (np.round(rng.randn(8, 8, 8) * 2) / 2, 0.0),              # values exactly on the isovalue (<=)
        "border_7": (np.pad(rng.randn(5, 5, 5), 1, constant_values=-1e6), 0.0),  # padded like extract_mesh
        "flat_5": (np.tile(np.array([1.0, 1.0, 0.0, 0.0, -1.0]), (5, 5, 1)), 0.0),  # f1 == f2 edges on the isovalue
        "empty_4": (np.ones((4, 4, 4)), 0.0),
    }
