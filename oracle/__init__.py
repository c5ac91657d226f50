"""CPU oracle for the AdaptivePnP_SCI hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + CPU PyTorch) of the reference's
two-stage online-adaptive ADMM reconstruction loop.  It exists to CHECK the
CUDA product in ``adaptivepnp_sci_b200/``; it is never shipped, never
imported by the product, and is executed only by

  * ``tests/``                      (as the checker),
  * ``__graft_entry__.smoke()``     (as the checker),
  * ``bench.py``                    (``cpu_baseline`` leg and ``--impl reference``).

Every function cites the reference ``file:line`` it restates (paths relative
to the reference tree).

Pinning status
--------------
* Bayer masks / Malvar-2004 (numpy variant): pinned by the reference's own
  doctest known-answer vectors (``tests/golden/colour_kats.json``).
* A/At, projections, Bayer remaps, tensor Malvar, FFDNet, FastDVDnet, the two
  adapters and both ADMM solvers: pinned against OUTPUTS OF THE REFERENCE
  ITSELF, imported from ``/root/reference`` in the build container by
  ``tests/golden/make_golden.py`` (which also asserts restatement ==
  reference); the resulting vectors are committed under ``tests/golden/``.
* ``denoise_tv_chambolle`` / PSNR / SSIM: the arithmetic lives in scikit-image
  0.18.1 (``readme.md:14``), which is neither vendored in the reference nor
  installed here -> **parity unpinned** for these three functions: the pin is
  this package's restatement of the published algorithm.
"""

BAYER = ((0, 0), (0, 1), (1, 0), (1, 1))  # RGGB, dvp_linear_inv_2_stage_ADMM_tensor_online.py:51
