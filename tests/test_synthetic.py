"""CPU: the product's synthetic generators equal the oracle's copy bit for bit."""
import numpy as np
import torch


def test_make_case_identical():
    from adaptivepnp_sci_b200 import synthetic as p
    from oracle import synthetic as o
    for bayer in (True, False):
        a, b = p.make_case(32, 48, 8, 7, bayer), o.make_case(32, 48, 8, 7, bayer)
        for x, y in zip(a, b):
            assert x.dtype == np.float32 and np.array_equal(x, y)
    meas, mask, orig = p.make_case(16, 16, 4, 1, True)
    assert np.array_equal(meas, (orig * mask).sum(2).astype(np.float32))
    assert set(np.unique(mask)) <= {0.0, 1.0}


def test_fastdvdnet_init_identical_and_loadable():
    from adaptivepnp_sci_b200 import synthetic as p
    from oracle import networks, synthetic as o
    a, b = p.fastdvdnet_synthetic_state_dict(), o.fastdvdnet_synthetic_state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k])
    networks.FastDVDnet().load_state_dict(a, strict=True)


def test_ddnet_init_identical_and_loadable():
    from adaptivepnp_sci_b200 import synthetic as p
    from oracle import networks, synthetic as o
    a, b = p.ddnet_synthetic_state_dict(), o.ddnet_synthetic_state_dict()
    assert sorted(a) == sorted(b)
    for k in a:
        assert torch.equal(a[k], b[k])
    networks.DDnet().load_state_dict(a, strict=True)
