"""Univariate g-and-k model (mirror of elfi/examples/gnk.py) + the 2-d order-statistic variant that
BASELINE config #5 needs (AdaptiveDistance over a (B, n_obs) summary matrix)."""
from functools import partial

import numpy as np
import scipy.stats as ss

from .. import device as dev
from .. import model as em
from .. import ops


def GNK(A, B, g, k, c=0.8, n_obs=50, batch_size=1, random_state=None):
    """Sample the g-and-k distribution through its quantile function (gnk.py:11-68);
    output shape (batch_size, n_obs, 1)."""
    A = np.asanyarray(A).reshape((-1, 1))
    B = np.asanyarray(B).reshape((-1, 1))
    g = np.asanyarray(g).reshape((-1, 1))
    k = np.asanyarray(k).reshape((-1, 1))
    z = ss.norm.rvs(size=(batch_size, n_obs), random_state=random_state)
    y = A + B * (1 + c * ((1 - np.exp(-g * z)) / (1 + np.exp(-g * z)))) * (1 + z**2)**k * z
    return y[:, :, np.newaxis]


def ss_order(y):
    """gnk.py:145-161, reproduced as is: np.sort over the LAST axis of (B, n_obs, 1), i.e. the
    identity."""
    return np.sort(dev.to_host(y))


def ss_sorted(y):
    """Order statistics per simulation, (B, n_obs): np.sort(y[:, :, 0], axis=1) on the device."""
    y = dev.to_host(y) if not dev.is_device_array(y) else y
    y2 = y[:, :, 0] if y.ndim == 3 else y
    return ops.rowsort(np.ascontiguousarray(y2) if isinstance(y2, np.ndarray) else y2.contiguous())


def euclidean_multiss(*simulated, observed):
    """gnk.py:115-142 (host callable over 3-d summaries)."""
    pts_sim = dev.to_host(simulated[0])
    pts_obs = dev.to_host(observed[0])
    d_ss_merged = np.sum((pts_sim - pts_obs)**2., axis=1)
    return np.sqrt(np.sum(d_ss_merged, axis=1))


def _base(n_obs, true_params, seed):
    m = em.new_model()
    if true_params is None:
        true_params = [3, 1, 2, .5]
    priors = [em.Prior('uniform', 0, 10, model=m, name=n) for n in ('A', 'B', 'g', 'k')]
    y_obs = GNK(*true_params, n_obs=n_obs, random_state=np.random.RandomState(seed))
    em.Simulator(partial(GNK, n_obs=n_obs), *priors, observed=y_obs, name='GNK')
    return m


def get_model(n_obs=50, true_params=None, seed=None):
    """Stock g-and-k task (gnk.py:71-112): ss_order + euclidean_multiss."""
    m = _base(n_obs, true_params, seed)
    default_ss = em.Summary(ss_order, m['GNK'], name='ss_order')
    em.Discrepancy(euclidean_multiss, default_ss, name='d')
    return m


def get_adaptive_model(n_obs=256, true_params=None, seed=None):
    """Config #5: (B, n_obs) order statistics + AdaptiveDistance (for AdaptiveDistanceSMC)."""
    m = _base(n_obs, true_params, seed)
    s = em.Summary(ss_sorted, m['GNK'], name='ss_sorted')
    em.AdaptiveDistance(s, name='d')
    return m


# ---------------------------------------------------------------------------- throughput mode
# Device-side priors, simulator and proposals (Philox streams; statistical parity with the host
# path).  The simulated (B, n_obs) matrix is born in HBM, sorted per row there (order statistics)
# and consumed by the nested-distance kernel; only accepted particles leave the device.
def gnk_device(A, B, g, k, c=0.8, n_obs=50, batch_size=1, random_state=None):
    """Device twin of GNK: returns a (batch_size, n_obs) CUDA tensor."""
    from .gauss import _key

    def as_dev(v):
        if dev.is_device_array(v):
            return v.reshape(-1)
        return dev.to_device(np.broadcast_to(np.asarray(v, dtype=np.float64).reshape(-1),
                                             (batch_size,)).copy())
    return ops.sim_gnk(as_dev(A), as_dev(B), as_dev(g), as_dev(k), n_obs=n_obs,
                       seed=_key(random_state), c=c)


class DeviceProposal:
    """SMC proposals / prior density on the device for the g-and-k model
    (pass an instance as ``device_proposal=`` to SMC / AdaptiveDistanceSMC)."""
    parameter_names = ['A', 'B', 'g', 'k']

    def __init__(self, lo=0.0, width=10.0):
        self.lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (4,)).copy()
        self.width = np.broadcast_to(np.asarray(width, dtype=np.float64), (4,)).copy()
        self.box = (list(self.lo), list(self.lo + self.width))

    def rvs(self, means, cov, weights, size, key, cdf=None):
        return ops.gm_rvs(means, cov, weights, size, seed=key, support=2, box=self.box, cdf=cdf)

    def logpdf(self, params):
        return ops.logprior_box(params, self.lo, self.width)


def get_device_model(n_obs=256, true_params=None, seed=None):
    """Config #5 in throughput mode: uniform(0, 10) priors drawn on the device, `gnk_device`,
    row-sorted order statistics and AdaptiveDistance.  Returns (model, DeviceProposal)."""
    from .gauss import _DeviceUniform
    if true_params is None:
        true_params = [3, 1, 2, .5]
    m = em.new_model()
    priors = [em.Prior(_DeviceUniform, 0, 10, model=m, name=n) for n in ('A', 'B', 'g', 'k')]
    y_obs = GNK(*true_params, n_obs=n_obs, random_state=np.random.RandomState(seed))
    em.Simulator(partial(gnk_device, n_obs=n_obs), *priors, observed=y_obs, name='GNK')
    s = em.Summary(ss_sorted, m['GNK'], name='ss_sorted')
    em.AdaptiveDistance(s, name='d')
    return m, DeviceProposal()
