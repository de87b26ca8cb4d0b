"""Numpy math utils — mirror of the reference's ``rtgs/utils/math.py``."""

import numpy as np


def sigmoid(x: np.ndarray) -> np.ndarray:
    """Sigmoid 1 / (1 + e^-x), evaluated in the dtype of ``x`` exactly as the reference does
    (utils/math.py:8-14), so float32 PLY columns give bit-identical activations."""
    return 1 / (1 + np.exp(-x))
