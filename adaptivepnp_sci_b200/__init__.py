"""adaptivepnp_sci_b200 — B200-native (sm_100a) hot path of AdaptivePnP_SCI.

Host side: Python mirrors of the reference's operator / solver / denoiser plug-in
interface (same names, argument meaning and error behaviour).  Device side: the
hand-written CUDA kernels in ``csrc/`` behind the C ABI of ``include/sci_b200.h``,
loaded with ctypes by ``_lib``.  Importing the package without the built library
raises ImportError: there is no CPU or PyTorch-op fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if libsci_b200.so is missing)

__version__ = "0.1.0"
