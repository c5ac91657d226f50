"""CPU: sci_host_legacy_normal reproduces numpy's legacy RandomState.normal bit for bit (values AND final
generator state), for even/odd counts, a pending cached gaussian, block boundaries and several thread counts."""
import numpy as np
import pytest


@pytest.mark.parametrize("n,threads", [(1, 1), (2, 1), (7, 4), (4096 * 9 + 1, 4), (300001, 16), (8 * 3 * 64 * 64, 8)])
def test_bit_exact_vs_numpy(n, threads):
    from adaptivepnp_sci_b200.fastdvdnet_adapter import fast_legacy_normal
    for seed, predraw in ((42, 0), (7, 1), (123, 623), (5, 2)):
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        if predraw:
            a.normal(size=predraw), b.normal(size=predraw)          # odd predraw leaves a cached gaussian pending
        ref = a.normal(0, 5 / 255, (n,))
        got = fast_legacy_normal(b, 0, 5 / 255, (n,), nthreads=threads)
        assert np.array_equal(ref, got)
        sa, sb = a.get_state(), b.get_state()
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]
        assert a.normal() == b.normal() and a.random_sample() == b.random_sample()   # streams continue identically


def test_two_consecutive_calls():
    from adaptivepnp_sci_b200.fastdvdnet_adapter import fast_legacy_normal
    a, b = np.random.RandomState(99), np.random.RandomState(99)
    for n in (5, 1000, 3, 77777):
        assert np.array_equal(a.normal(1.5, 2.0, (n,)), fast_legacy_normal(b, 1.5, 2.0, (n,), nthreads=3))
