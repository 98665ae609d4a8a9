"""Minimal stand-in for the two scikit-image helpers the reference imports (utils/util_image.py:14), so that the
reference tree can be imported in this container by the oracle / golden tooling and the pinning tests.  Test
infrastructure only."""
import numpy as np


def img_as_ubyte(x):
    x = np.asarray(x)
    if x.dtype == np.uint8:
        return x
    return np.clip(np.rint(x.astype(np.float64) * 255.0), 0, 255).astype(np.uint8)


def img_as_float32(x):
    x = np.asarray(x)
    if x.dtype == np.uint8:
        return x.astype(np.float32) / 255.0
    return x.astype(np.float32)
