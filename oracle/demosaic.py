"""Oracle: Malvar-2004 demosaicing, both variants found in the reference.

Test infrastructure.  ``malvar2004_numpy`` restates the vendored colour-science
function (scipy ``reflect`` boundary) and is pinned by the doctest known-answer
vectors at packages/colour_demosaicing/bayer/demosaicing/malvar2004.py:70-95.
``malvar2004_tensor`` restates the ``_tensor`` variant the hot path actually
calls (malvar2004.py:169-246: torch ``reflect`` padding = mirror WITHOUT edge
repeat, cross-correlation, filters built in float64 then cast to float32) and
is pinned against the reference function itself by tests/golden/make_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.ndimage import convolve

_GR_GB = np.array([[0, 0, -1, 0, 0],
                   [0, 0, 2, 0, 0],
                   [-1, 2, 4, 2, -1],
                   [0, 0, 2, 0, 0],
                   [0, 0, -1, 0, 0]], np.float64) / 8
_Rg_RB = np.array([[0, 0, 0.5, 0, 0],
                   [0, -1, 0, -1, 0],
                   [-1, 4, 5, 4, -1],
                   [0, -1, 0, -1, 0],
                   [0, 0, 0.5, 0, 0]], np.float64) / 8
_Rg_BR = _Rg_RB.T.copy()
_Rb_BB = np.array([[0, 0, -1.5, 0, 0],
                   [0, 2, 0, 2, 0],
                   [-1.5, 0, 6, 0, -1.5],
                   [0, 2, 0, 2, 0],
                   [0, 0, -1.5, 0, 0]], np.float64) / 8


def masks_CFA_Bayer(shape, pattern="RGGB"):
    """packages/colour_demosaicing/bayer/masks.py:23-72."""
    ch = {c: np.zeros(shape, bool) for c in "RGB"}
    for c, (y, x) in zip(pattern.upper(), [(0, 0), (0, 1), (1, 0), (1, 1)]):
        ch[c][y::2, x::2] = True
    return ch["R"], ch["G"], ch["B"]


def _select(CFA, R_m, G_m, B_m, G_f, RB_row, RB_col, RB_diag, where, rows_any, cols_any):
    R = CFA * R_m
    G = CFA * G_m
    Bc = CFA * B_m
    G = where(R_m | B_m, G_f, G)
    R_r, R_c = rows_any(R_m), cols_any(R_m)
    B_r, B_c = rows_any(B_m), cols_any(B_m)
    R = where(R_r & B_c, RB_row, R)
    R = where(B_r & R_c, RB_col, R)
    Bc = where(B_r & R_c, RB_row, Bc)
    Bc = where(R_r & B_c, RB_col, Bc)
    R = where(B_r & B_c, RB_diag, R)
    Bc = where(R_r & R_c, RB_diag, Bc)
    return R, G, Bc


def malvar2004_numpy(CFA, pattern="RGGB"):
    """malvar2004.py:38-160 (numpy / scipy.ndimage.convolve, mode='reflect')."""
    CFA = np.asarray(CFA, np.float64)
    R_m, G_m, B_m = masks_CFA_Bayer(CFA.shape, pattern)
    R, G, Bc = _select(
        CFA, R_m, G_m, B_m,
        convolve(CFA, _GR_GB), convolve(CFA, _Rg_RB), convolve(CFA, _Rg_BR), convolve(CFA, _Rb_BB),
        np.where,
        lambda m: np.any(m, axis=1)[:, None] & np.ones(m.shape, bool),
        lambda m: np.any(m, axis=0)[None, :] & np.ones(m.shape, bool))
    return np.stack([R, G, Bc], axis=-1)


def malvar2004_tensor(CFA, R_m, G_m, B_m):
    """malvar2004.py:169-246 — the variant on the hot path. CFA [H,W] float32
    CPU tensor, boolean masks from ``masks_CFA_Bayer_tensor``; returns [H,W,3]."""
    k = [torch.tensor(a, dtype=torch.float64).float()[None, None] for a in (_GR_GB, _Rg_RB, _Rg_BR, _Rb_BB)]
    pad = F.pad(CFA[None, None], (2, 2, 2, 2), mode="reflect")
    G_f, RB_row, RB_col, RB_diag = [F.conv2d(pad, kk)[0, 0] for kk in k]
    R, G, Bc = _select(
        CFA, R_m, G_m, B_m, G_f, RB_row, RB_col, RB_diag,
        torch.where,
        lambda m: torch.any(m, dim=1)[:, None] & torch.ones(m.shape, dtype=torch.bool),
        lambda m: torch.any(m, dim=0)[None, :] & torch.ones(m.shape, dtype=torch.bool))
    return torch.stack([R, G, Bc], dim=-1)
