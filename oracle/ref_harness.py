"""Import the REAL reference (read-only tree at /root/reference) on CPU.

Test infrastructure, BUILD-CONTAINER ONLY: /root/reference does not exist on
the GPU box, so nothing under ``tests/ -m gpu``, ``smoke()`` or ``bench.py``
may call this.  It is used by ``tests/golden/make_golden.py`` to (a) generate
the committed golden vectors from the reference's own functions and (b) assert
that the restatements in this package reproduce them.

What it does:
  * stubs the imports that are missing here but numerically unused on the hot
    path (imageio, h5py, matplotlib.pyplot, tensorboardX, colour.utilities,
    cv2/torchvision are real);
  * provides ``skimage`` backed by this package's restatements of
    ``denoise_tv_chambolle`` / PSNR / SSIM (the only third-party arithmetic on
    the path that is absent -> those three stay "parity unpinned");
  * neutralises the hard-coded ``.cuda()`` calls so the reference runs on CPU.
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SCI_REFERENCE_ROOT", "/root/reference")
# ``baseline/_ref``: a git-ignored copy of the reference's Python sources made by ``baseline/install_ref.py`` in the build
# container.  Unlike /root/reference it travels to the GPU box with the snapshot, where ``bench.py`` runs the UNMODIFIED
# reference on the same B200 through PyTorch eager (``gpu_eager_baseline``) - the honest same-box comparator.
BASELINE_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
_loaded = {}


def available(root=None):
    root = root or REF_ROOT
    return os.path.isdir(root) and os.path.isfile(os.path.join(root, "utilspy.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_shims(cpu=True):
    import torch
    from . import iqa, tv_chambolle

    for name in ("imageio", "h5py"):
        if name not in sys.modules:
            _stub(name)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot", xlim=None)
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=object)
    if "colour" not in sys.modules:
        col = _stub("colour")
        col.utilities = _stub(
            "colour.utilities",
            as_float_array=lambda a, dtype=None: np.asarray(a, dtype=np.float64),
            tstack=lambda a: np.stack([np.asarray(x) for x in a], axis=-1),
            tsplit=lambda a: [np.asarray(a)[..., i] for i in range(np.asarray(a).shape[-1])],
            ANCILLARY_COLOUR_SCIENCE_PACKAGES={})
    if "skimage" not in sys.modules:
        sk = _stub("skimage", __version__="0.18.1")
        sk.restoration = _stub("skimage.restoration", denoise_tv_chambolle=tv_chambolle.denoise_tv_chambolle)
        sk.metrics = _stub("skimage.metrics", peak_signal_noise_ratio=iqa.compare_psnr,
                           structural_similarity=iqa.compare_ssim)
        sk.color = _stub("skimage.color")
        sk.io = _stub("skimage.io")
        sk.metrics.simple_metrics = _stub("skimage.metrics.simple_metrics",
                                          peak_signal_noise_ratio=iqa.compare_psnr)
        sk.metrics._structural_similarity = _stub("skimage.metrics._structural_similarity",
                                                  structural_similarity=iqa.compare_ssim)
        sk.measure = _stub("skimage.measure", compare_psnr=iqa.compare_psnr, compare_ssim=iqa.compare_ssim)
    if not cpu:
        return
    # CPU execution of hard-coded .cuda() calls
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    torch.cuda.manual_seed = lambda *a, **k: None
    _type = torch.Tensor.type
    # test_fastdvdnet.py:214 / test_ffdnet_ipol.py:164 ask for 'torch.cuda.FloatTensor' by name
    torch.Tensor.type = lambda self, dtype=None, *a, **k: _type(
        self, "torch.FloatTensor" if (dtype == "torch.cuda.FloatTensor" or dtype is getattr(torch.cuda, "FloatTensor", None))
        else dtype, *a, **k)


def load(root=None, device="cpu"):
    """Returns a namespace with the reference's hot-path modules.  ``device='cuda'`` leaves the reference's hard-coded
    ``.cuda()`` calls alone (the GPU-eager baseline of bench.py); the default neutralises them (CPU golden generation)."""
    if _loaded:
        return _loaded["ns"]
    root = root or REF_ROOT
    if not available(root):
        raise RuntimeError("reference tree not present at %s" % root)
    _install_shims(cpu=(device == "cpu"))
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    ns = types.SimpleNamespace()
    ns.utilspy = importlib.import_module("utilspy")
    ns.dvp = importlib.import_module("dvp_linear_inv_2_stage_ADMM_tensor_online")
    ns.utils_image = importlib.import_module("utils.utils_image")
    ns.malvar = importlib.import_module("packages.colour_demosaicing.bayer.demosaicing.malvar2004")
    ns.masks = importlib.import_module("packages.colour_demosaicing.bayer.masks")
    ns.network_ffdnet = importlib.import_module("models.network_ffdnet")
    ns.fastdvd_models = importlib.import_module("packages.fastdvdnet.models")
    ns.fastdvd_adapter = importlib.import_module("packages.fastdvdnet.test_fastdvdnet")
    ns.fastdvd_driver = importlib.import_module("packages.fastdvdnet.fastdvdnet")
    ns.ffdnet_adapter = importlib.import_module("packages.ffdnet.test_ffdnet_ipol")
    ns.ffdnet_ipol_models = importlib.import_module("packages.ffdnet.models")
    ns.network_demosaicking = importlib.import_module("models.network_demosaicking")
    ns.ddnet_adapter = importlib.import_module("packages.DDnet.DDnet_test")
    import torch
    torch.autograd.set_detect_anomaly(False)   # test_ffdnet_ipol.py:26 turns it on globally (perf only)
    _loaded["ns"] = ns
    return ns
