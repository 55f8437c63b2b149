"""Array-level operators of the hot path: thin Python wrappers over the C ABI.

Each wrapper accepts CUDA tensors (or host arrays, which are uploaded) and returns CUDA
tensors.  Reference call sites are cited per function (paths relative to the reference repo).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import device as dev

MAX_NESTED = 32


def _matrix(S):
    if dev.is_device_array(S) and S.dtype == torch.float64 and S.dim() == 2 \
            and S.stride(1) == 1 and S.stride(0) >= S.shape[1]:
        return S  # row-strided views are consumed in place (leading dimension = stride(0))
    S = dev.to_device(S)
    if S.dim() == 1:
        S = S[:, None]
    if S.dim() != 2:
        raise ValueError('expected a 2-d (batch, dim) array, got shape {}'.format(tuple(S.shape)))
    if S.stride(1) != 1:
        S = S.contiguous()
    return S


def dist_euclid(S, obs, w=None, thresholds=None, want_indices=True):
    """Euclidean / nested weighted distances of the rows of S to ``obs`` + acceptance.

    Replaces ``cdist(S, obs, 'euclidean'[, w=w])`` reached through
    elfi/model/utils.py:37-52 and elfi/model/elfi_model.py:1037, 1135-1151, and the mask of
    elfi/methods/inference/samplers.py:223-225.

    Parameters
    ----------
    S : (B, D) array
    obs : (D,) or (1, D) array
    w : None, (D,) or (K, D) -- cdist's ``w`` per nested column (a row of ones == unweighted)
    thresholds : None, float or (K,) -- accept rows with all_k(d[:, k] <= thresholds[k])

    Returns
    -------
    d : (B,) tensor if K == 1 and w was not 2-d, else (B, K)
    acc_idx : int32 tensor of accepted row indices, ascending (None without thresholds)
    """
    S = _matrix(S)
    B, D = S.shape
    obs_t = dev.to_device(obs).reshape(-1)
    if obs_t.numel() != D:
        raise ValueError('XA and XB must have the same number of columns '
                         '(i.e. feature dimension.)')
    squeeze = True
    W = None
    K = 1
    if w is not None:
        W = dev.to_device(w)
        if W.dim() == 1:
            W = W[None, :]
        else:
            squeeze = False
        K = W.shape[0]
        if W.shape[1] != D:
            raise ValueError('weights must have {} columns'.format(D))
        if K > MAX_NESTED:
            raise ValueError('at most {} nested distances are supported'.format(MAX_NESTED))
    thr = None
    if thresholds is not None:
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if thr.shape[0] != K:
            raise ValueError('need one threshold per distance column ({} != {})'.format(
                thr.shape[0], K))
    d = dev.empty((B, K))
    acc_idx = n_acc = None
    if thr is not None:
        n_acc = torch.zeros(1, dtype=torch.int64, device='cuda')
        if want_indices:
            acc_idx = dev.empty((max(B, 1),), dtype=torch.int32)
    _lib.call('elfi_b200_dist_euclid_thr_f64', dev.context(), dev.ptr(S), S.stride(0), B, D,
              dev.ptr(obs_t), dev.ptr(W), K, dev.ptr(thr), dev.ptr(d), dev.ptr(acc_idx),
              dev.ptr(n_acc), dev.stream_ptr())
    if thr is not None:
        n = int(n_acc.item())
        if want_indices:
            acc_idx = acc_idx[:n]
        else:
            acc_idx = n
    if squeeze:
        d = d.reshape(B)
    return d, acc_idx


def dist_euclid_host(S, obs, w=None, thresholds=None, return_distances=True):
    """Host-buffer variant (elfi_b200_dist_euclid_thr_f64_host): numpy in, numpy out."""
    S = np.asarray(S, dtype=np.float64)
    if S.ndim == 1:
        S = S[:, None]
    if S.strides[1] != 8:
        S = np.ascontiguousarray(S)
    B, D = S.shape
    ld = S.strides[0] // 8 if B > 1 else D
    obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
    K = 1
    W = None
    if w is not None:
        W = np.ascontiguousarray(np.atleast_2d(w), dtype=np.float64)
        K = W.shape[0]
    thr = None if thresholds is None else np.ascontiguousarray(np.atleast_1d(thresholds),
                                                               dtype=np.float64)
    d = np.empty((B, K)) if return_distances else None
    idx = np.empty(max(B, 1), dtype=np.int32) if thr is not None else None
    n = ctypes.c_int64(0)
    _lib.call('elfi_b200_dist_euclid_thr_f64_host', dev.context(), dev.ptr(S), ld, B, D,
              dev.ptr(obs), dev.ptr(W), K, dev.ptr(thr), dev.ptr(d), dev.ptr(idx),
              ctypes.byref(n) if thr is not None else None)
    if d is not None and K == 1 and (w is None or np.ndim(w) == 1):
        d = d.reshape(B)
    return d, (idx[:n.value] if idx is not None else None)
