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


@pytest.mark.parametrize("chunk_blocks", [2, 3, 7, 64])
def test_chunk_boundaries(monkeypatch, chunk_blocks):
    """The word stream is produced chunk by chunk (SCI_RNG_CHUNK_BLOCKS state blocks per chunk) while worker threads consume the
    previous chunk: tiny chunks put candidates across every kind of chunk / block boundary, for every start position class."""
    from adaptivepnp_sci_b200.fastdvdnet_adapter import fast_legacy_normal
    monkeypatch.setenv("SCI_RNG_CHUNK_BLOCKS", str(chunk_blocks))
    for seed, predraw, n, threads in ((1, 0, 20011, 3), (2, 1, 4999, 2), (3, 311, 12345, 5), (4, 312, 7000, 1), (5, 2, 1247, 4),
                                      (6, 0, 156 * 2 * chunk_blocks, 2)):
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        if predraw:
            a.random_sample(predraw), b.random_sample(predraw)      # 2 words each: moves the start position inside the block
        ref = a.normal(0.25, 1.5, (n,))
        got = fast_legacy_normal(b, 0.25, 1.5, (n,), nthreads=threads)
        assert np.array_equal(ref, got)
        sa, sb = a.get_state(), b.get_state()
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]
        assert np.array_equal(a.normal(size=5), b.normal(size=5))


def test_large_draw_streams_in_constant_memory():
    """4 M normals = ~10 M words, several default-size chunks: equal to numpy, state included."""
    from adaptivepnp_sci_b200.fastdvdnet_adapter import fast_legacy_normal
    a, b = np.random.RandomState(2024), np.random.RandomState(2024)
    n = 4_000_001
    assert np.array_equal(a.normal(0, 5 / 255, (n,)), fast_legacy_normal(b, 0, 5 / 255, (n,), nthreads=4))
    assert np.array_equal(a.get_state()[1], b.get_state()[1]) and a.get_state()[2:] == b.get_state()[2:]
