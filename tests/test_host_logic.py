"""CPU: small pieces of host logic around the hot path (no GPU, no compute calls into the library)."""
import os

import numpy as np


def test_rng_helper_threads_share_the_node(monkeypatch):
    """The host-noise helper leaves cores for the launch threads of all ranks of the node (8-GPU scaling fix)."""
    from adaptivepnp_sci_b200 import fastdvdnet_adapter as fa
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert fa._default_rng_threads() == max(1, min(8, cores - 1))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert fa._default_rng_threads() == max(1, min(8, cores // 8 - 1))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", str(4 * cores))
    assert fa._default_rng_threads() == 1


def test_warm_start_handoff_roundtrip(tmp_path):
    """Stage 1 -> stage 2 hand-off file (results/savedmat/_Admm_tv_<name>8.mat, key v_Admm_tv_denoise; reference
    ADMM_TV_Warm_Start_save.py:174-178 <-> two_stage_ADMM_Online_FFD_Warm.py:171-176)."""
    from adaptivepnp_sci_b200 import matio
    v = np.random.default_rng(0).random((16, 24, 16)).astype(np.float32)
    d = str(tmp_path) + "/"
    p = matio.save_warm_start(d, "Beauty_bayer", 8, v, np.ones(16, np.float32), np.ones(16, np.float32))
    assert os.path.basename(p) == "_Admm_tv_Beauty_bayer8.mat"
    back = matio.load_warm_start(d, "Beauty_bayer", 8)
    assert back.dtype == np.float32 and np.array_equal(back, v)


def test_synthetic_dataset_fallback_shapes():
    """Without dataset files the loaders fall back to the deterministic synthetic videos (scale 0..255 like the .mat data)."""
    from adaptivepnp_sci_b200 import matio
    meas, mask, orig = matio.load_video("/nonexistent", "Jockey_bayer", nmea=2, synthetic_shape=(32, 48, 8), force_synthetic=True)
    assert meas.shape == (32, 48, 2) and mask.shape == (32, 48, 8) and orig.shape == (32, 48, 16)
    assert meas.dtype == np.float32 and float(orig.max()) <= 255.0 and float(orig.max()) > 1.0
    assert np.allclose(meas[:, :, 1], (orig[:, :, 8:] * mask).sum(2), rtol=1e-5)


def test_script_schedules_cover_all_videos():
    from adaptivepnp_sci_b200 import matio, stage2_script as s
    for table in (s.FFD_TABLE, s.FASTDVD_TABLE, s.FFD_DEEP, s.FASTDVD_DEEP):
        assert set(table) == set(matio.VIDEOS)
    for name, (sig, iters, lr, upi, interval, times) in s.FASTDVD_TABLE.items():
        assert len(sig) == len(iters) and lr in (2e-6, 2e-7) and upi in (1, 2)
    for name, (sig, iters, interval) in s.FASTDVD_DEEP.items():
        assert len(sig) == len(iters) and interval >= 1


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm of the contract) emits exactly one JSON line on stdout with the agreed keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1",
                        "--warmup", "0"],        # config 1 (TV, 256x256x8) is the CPU-cheap one; the headline config takes minutes on 8 cores
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "admm_iters_per_sec" and d["unit"] == "iters/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_count_updates_and_noise_skip():
    """Host logic of the video sharding: the fine-tune noise stream is fast-forwarded exactly."""
    from adaptivepnp_sci_b200 import stage2_script as s
    assert s.count_updates([21, 2], 9, -1) == 2 and s.count_updates([18], 9, 1) == 1 and s.count_updates([6, 6, 4], 6, -1) == 2
    np.random.seed(42)
    a = [np.random.normal(0, 5 / 255, (2, 3, 8, 8)) for _ in range(3)]
    np.random.seed(42)
    s.skip_finetune_noise(2, (2, 3, 8, 8))
    b = np.random.normal(0, 5 / 255, (2, 3, 8, 8))
    assert np.array_equal(a[2], b)


def test_matlab_v73_reader_roundtrip(tmp_path):
    """The HDF5 subset of MATLAB -v7.3 files (user block, v0 superblock, old-style group, v1 object headers, contiguous and
    chunked + shuffle + deflate layouts) through the package's own reader, and the dataset loader on top of it
    (ADMM_TV_Warm_Start_save.py:69-93: same keys, same transposes).  No third-party HDF5 writer exists in this image: the
    file is produced by the module's spec-following test writer."""
    from adaptivepnp_sci_b200 import h5lite, matio
    rng = np.random.default_rng(0)
    H, W, B, nmea = 48, 32, 8, 4
    arrs = {"mask_bayer": (rng.random((B, W, H)) > 0.5).astype(np.uint8),            # MATLAB stores [H,W,B] column-major
            "meas_bayer": rng.random((nmea, W, H)) * 255,                             # -> HDF5 sees the reversed shape
            "orig_bayer": (rng.random((B * nmea, W, H)) * 255).astype(np.float32),
            "orig": rng.integers(0, 255, (B * nmea, 3, W, H)).astype(np.uint8)}
    d = tmp_path / "cacti"
    d.mkdir()
    h5lite.write_mat73(str(d / "Beauty_bayer.mat"), arrs, chunks={"orig_bayer": (5, 16, 32), "orig": (7, 3, 10, 9)})
    f = h5lite.File(str(d / "Beauty_bayer.mat"))
    assert sorted(f.keys()) == sorted(arrs) and "orig" in f
    for k, v in arrs.items():
        r = f[k]
        assert r.shape == v.shape and r.dtype == v.dtype and np.array_equal(r, v), k
    meas, mask, orig, orig_real = matio.load_video(str(d), "Beauty_bayer", nmea, with_orig_real=True)
    assert meas.shape == (H, W, nmea) and mask.shape == (H, W, B) and orig.shape == (H, W, B * nmea)
    assert np.array_equal(mask, np.float32(arrs["mask_bayer"]).transpose(2, 1, 0))
    assert np.array_equal(orig_real, arrs["orig"]) and matio.video_shape(str(d), "Beauty_bayer") == (H, W, B)
    with __import__("pytest").raises(ValueError):
        (d / "bad.mat").write_bytes(b"MATLAB 5.0 MAT-file" + b"\0" * 2000)
        h5lite.File(str(d / "bad.mat"))


def test_gray_fastdvdnet_embedding_is_exact():
    """The single-channel FastDVDnet runs on the colour engine through a zero embedding (fastdvdnet_models.FastDVDnet,
    num_color_channels=1).  CPU check of the claim itself: the colour-shaped twin's weights, loaded into the ORACLE colour
    network, reproduce the oracle gray network on the first plane (the other planes stay zero), and the twin follows in-place
    weight changes without entering the gray model's state_dict."""
    import torch
    from adaptivepnp_sci_b200 import synthetic as syn
    from adaptivepnp_sci_b200.fastdvdnet_models import FastDVDnet
    from oracle import networks, synthetic
    sd = syn.fastdvdnet_gray_synthetic_state_dict()
    assert all(torch.equal(sd[k], v) for k, v in synthetic.fastdvdnet_gray_synthetic_state_dict().items())
    m = FastDVDnet(num_input_frames=5, num_color_channels=1)
    m.load_state_dict(sd, strict=True)
    m.eval()
    twin = m._colour_twin()
    assert list(m.state_dict().keys()) == list(sd.keys())               # the twin is not a sub-module
    oc = networks.FastDVDnet(5, 3); oc.load_state_dict(twin.state_dict(), strict=True); oc.eval()
    og = networks.FastDVDnet(5, 1); og.load_state_dict(sd, strict=True); og.eval()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 5, 24, 32, generator=g)
    x3 = torch.zeros(1, 15, 24, 32); x3[:, 0::3] = x
    nm = torch.full((1, 1, 24, 32), 12 / 255)
    with torch.no_grad():
        want, got = og(x, nm), oc(x3, nm)
    assert float((got[:, 0:1] - want).abs().max()) < 1e-6 and float(got[:, 1:].abs().max()) == 0.0
    with torch.no_grad():
        m.temp1.inc.convblock[0].weight.mul_(1.5)
    w_gray, w_twin = m.temp1.inc.convblock[0].weight.detach(), m._colour_twin().temp1.inc.convblock[0].weight.detach()
    assert torch.equal(w_twin[:, 0], w_gray[:, 0]) and torch.equal(w_twin[:, 3], w_gray[:, 1]) and float(w_twin[:, 1:3].abs().max()) == 0.0


def test_ipol_ffdnet_first_layer_reordering():
    """IPOL-flavour FFDNet on the FFDNet engine: the engine's input kernel lays the down-sampled stack out as
    [4C sub-images | sigma]; the module hands it a first convolution with its input columns in that order (the C noise
    columns, which all see sigma, added up).  CPU check against the oracle network with the reference's [noise | sub-images] order."""
    import torch
    import torch.nn.functional as F
    from adaptivepnp_sci_b200 import synthetic as syn
    from adaptivepnp_sci_b200.ffdnet_ipol_models import FFDNet
    from oracle import networks
    for C in (1, 3):
        sd = syn.ffdnet_ipol_synthetic_state_dict(C)
        m = FFDNet(C); m.load_state_dict(sd, strict=True); m.eval()
        layers = m.conv_layers()
        assert len(layers) == (15 if C == 1 else 12) and layers[0][1] is None and layers[1][1] is not None and layers[-1][1] is None
        o = networks.FFDNetIPOL(C); o.load_state_dict(sd, strict=True); o.eval()
        g = torch.Generator().manual_seed(C)
        down = torch.rand(2, 4 * C, 10, 12, generator=g)                 # the 4C sub-images of some frame
        sigma = 20 / 255
        with torch.no_grad():
            ours = F.conv2d(torch.cat((down, torch.full((2, 1, 10, 12), sigma)), 1), layers[0][0].weight, padding=1)
            ref = F.conv2d(torch.cat((torch.full((2, C, 10, 12), sigma), down), 1), o.intermediate_dncnn.itermediate_dncnn[0].weight, padding=1)
        assert float((ours - ref).abs().max()) < 1e-5
    import pytest
    m.train()
    with pytest.raises(NotImplementedError):
        m.engine()


def test_matlab_v73_reader_fuzz(tmp_path):
    """Property test of the HDF5-subset reader against its spec-following writer: random ranks (1-4), extents, element types,
    contiguous / chunked (+ shuffle + deflate) layouts with ragged edge chunks, user-block sizes."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    from adaptivepnp_sci_b200 import h5lite
    dtypes = st.sampled_from([np.uint8, np.int8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.float32, np.float64])

    @st.composite
    def arrays(draw):
        shape = tuple(draw(st.integers(1, 9)) for _ in range(draw(st.integers(1, 4))))
        a = (np.random.default_rng(draw(st.integers(0, 1000))).random(shape) * 200 - 50).astype(draw(dtypes))
        return a, (tuple(draw(st.integers(1, s)) for s in shape) if draw(st.booleans()) else None)

    count = [0]

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(st.lists(arrays(), min_size=1, max_size=4), st.sampled_from([0, 512, 1024]))
    def roundtrip(items, userblock):
        count[0] += 1
        arrs = {"v%d" % i: a for i, (a, c) in enumerate(items)}
        chunks = {"v%d" % i: c for i, (a, c) in enumerate(items) if c is not None}
        p = str(tmp_path / ("t%d.mat" % count[0]))
        h5lite.write_mat73(p, arrs, chunks=chunks, userblock=userblock)
        f = h5lite.File(p)
        assert sorted(f.keys()) == sorted(arrs)
        for k, v in arrs.items():
            r = f[k]
            assert r.shape == v.shape and r.dtype == v.dtype and np.array_equal(r, v), (k, v.shape, v.dtype, chunks.get(k))

    roundtrip()
