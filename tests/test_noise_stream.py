"""CPU: the run-ahead fine-tune noise producer is EXACTLY equivalent to the reference's inline
np.random.normal draws (values and global RNG state), also across reseeds and foreign draws."""
import time

import numpy as np


def test_noise_stream_exact_semantics():
    from adaptivepnp_sci_b200.fastdvdnet_adapter import NoiseStream
    shape = (2, 3, 16, 16)
    np.random.seed(42)
    ref = [np.random.normal(0, 5 / 255, shape) for _ in range(5)]
    ref_next = np.random.rand()
    ns = NoiseStream()
    np.random.seed(42)
    ns.prefetch(shape)
    time.sleep(0.05)                                   # let the helper run ahead
    got = [ns.get(shape) for _ in range(5)]
    assert all(np.array_equal(a, b) for a, b in zip(ref, got)) and np.random.rand() == ref_next
    # reseed in between (the scripts reseed once per run; tests reseed per case)
    np.random.seed(7)
    r2 = [np.random.normal(0, 5 / 255, shape) for _ in range(2)]
    x2 = np.random.rand()
    np.random.seed(7)
    g2 = [ns.get(shape) for _ in range(2)]
    assert all(np.array_equal(a, b) for a, b in zip(r2, g2)) and np.random.rand() == x2
    # somebody else draws from the global RNG between two fine-tune calls
    np.random.seed(9)
    a1, f, a2 = np.random.normal(0, 5 / 255, shape), np.random.rand(3), np.random.normal(0, 5 / 255, shape)
    np.random.seed(9)
    b1 = ns.get(shape)
    f2 = np.random.rand(3)
    b2 = ns.get(shape)
    assert np.array_equal(a1, b1) and np.array_equal(f, f2) and np.array_equal(a2, b2)
    # shape change
    np.random.seed(3)
    c = np.random.normal(0, 5 / 255, (1, 3, 8, 8))
    np.random.seed(3)
    assert np.array_equal(ns.get((1, 3, 8, 8)), c)
