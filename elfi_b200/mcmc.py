"""MCMC samplers and chain diagnostics used by BOLFI.sample (elfi/methods/mcmc.py).

`nuts` is the No-U-Turn sampler with dual-averaging step size adaptation (Hoffman & Gelman 2014,
Algorithm 6) and `metropolis` a Gaussian random walk; both consume their RandomState in the same
order as the reference (mcmc.py:105-429), so a chain started from the same seed visits the same
points -- tests/golden/mcmc_*.npz holds chains produced by the reference on analytic targets.
The targets are host callables; in BOLFI they are BolfiPosterior.logpdf / gradient_logpdf, whose
GP mean, variance and gradients come from one device call per point (bo.py).
"""
import logging
from collections import namedtuple

import numpy as np

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------- diagnostics
def _pooled_variance(chains):
    """Within / between chain variances of an (M, N) array -> (var_within, var_pooled)."""
    n = chains.shape[1]
    within = np.mean(np.var(chains, ddof=1, axis=1))
    between = n * np.var(np.mean(chains, axis=1), ddof=1) if chains.shape[0] > 1 else 0
    return within, ((n - 1.) * within + between) / n


def eff_sample_size(chains):
    """Effective sample size of one or more chains of one parameter (mcmc.py:11-60; BDA3,
    Stan manual): N M / (1 + 2 sum_t rho_t) with the multi-chain autocorrelation rho_t from the
    variogram, truncated at the first negative estimate."""
    chains = np.atleast_2d(chains)
    n_chains, n_samples = chains.shape
    means = np.mean(chains, axis=1)
    var_within, var_pooled = _pooled_variance(chains)
    # autocovariance at every lag through the FFT (zero padded to avoid wrap-around)
    n_fft = int(2 ** np.ceil(1 + np.log2(n_samples)))
    spectrum = np.abs(np.fft.rfft(chains - means[:, None], n_fft)) ** 2
    autocov = np.fft.irfft(spectrum)[:, :n_samples].real / np.arange(n_samples, 0, -1)
    rho_sum = 0.
    for lag in range(1, n_samples):
        rho = 1. - (var_within - np.mean(autocov[:, lag])) / var_pooled
        if not rho >= 0:
            break
        rho_sum += rho
    return n_chains * n_samples / (1. + 2. * rho_sum)


def gelman_rubin_statistic(chains):
    """Split-chain potential scale reduction factor R-hat (mcmc.py:63-102)."""
    chains = np.atleast_2d(chains)
    n_chains, n_samples = chains.shape
    half = n_samples // 2
    halves = chains[:, :2 * half].reshape((2 * n_chains, half))
    var_within, var_pooled = _pooled_variance(halves)
    return np.sqrt(var_pooled / var_within)


# -------------------------------------------------------------------------------------- NUTS
# The samplers are written as coroutines: wherever the log density or its gradient is needed the
# chain yields a request ('logpdf' | 'grad', point) and is resumed with the value.  `nuts` /
# `metropolis` drive one chain with plain callables (the reference's interface); `run_lockstep`
# advances many chains together and answers all their pending requests with ONE batched
# evaluation -- on the device that is one gp_predict_grad launch for all chains instead of one
# per chain and point.  A chain owns its RandomState, so its path does not depend on the driver.

# One (sub)tree of the doubling procedure: its two ends, the candidate drawn from it, the number of
# points inside the slice, whether it may be extended, and the Metropolis statistics that drive the
# step size adaptation.
_Tree = namedtuple('_Tree', 'left p_left right p_right candidate n_slice ok mh_sum n_steps '
                            'diverged outside')

_MAX_ENERGY_ERROR = 1000.   # a leaf whose joint falls this far below the slice level has diverged
LOGPDF, GRAD = 'logpdf', 'grad'


def _no_u_turn(tree_left, p_left, tree_right, p_right):
    span = tree_right - tree_left
    return np.inner(span, p_left) >= 0 and np.inner(span, p_right) >= 0


class _Nuts:
    """Tree building of one chain; every method that needs the target is a coroutine."""

    def __init__(self, random_state):
        self.rs = random_state

    def leapfrog(self, x, p, step):
        p_half = p + 0.5 * step * (yield (GRAD, x))
        x_new = x + step * p_half
        p_new = p_half + 0.5 * step * (yield (GRAD, x_new))
        return x_new, p_new

    # ---- initial step size: double / halve until the acceptance probability crosses 1/2 ------
    def initial_stepsize(self, x0, target0, max_retry_inits):
        grad0 = yield (GRAD, x0)
        logger.debug("NUTS: Trying to find initial stepsize from point {} with gradient {}."
                     .format(x0, grad0))

        def trial(p0, step):
            p1 = p0 + 0.5 * step * grad0
            x1 = x0 + step * p1
            p1 = p1 + 0.5 * step * (yield (GRAD, x1))
            return (yield (LOGPDF, x1)) - 0.5 * np.inner(p1, p1)

        for attempt in range(max_retry_inits):     # may step outside the prior support
            step = np.exp(-attempt)
            p0 = self.rs.randn(*x0.shape)
            joint1 = yield from trial(p0, step)
            joint0 = target0 - 0.5 * np.inner(p0, p0)
            if np.isfinite(joint1):
                break
            if attempt == max_retry_inits - 1:
                raise ValueError(
                    "NUTS: Cannot find acceptable stepsize starting from point {}. All "
                    "trials ended in region with 0 probability.".format(x0))
            logger.debug("NUTS: Problem finding acceptable stepsize, now {}. Retrying {}/{}."
                         .format(step, attempt + 1, max_retry_inits))
        sign = 1 if np.exp(joint1 - joint0) > 0.5 else -1
        factor = 2. if sign == 1 else 0.5
        while factor * np.exp(sign * (joint1 - joint0)) > 1.:
            step *= factor
            if step == 0. or step > 1e7:      # bounds as in Stan
                raise SystemExit("NUTS: Found invalid stepsize {} starting from point {}."
                                 .format(step, x0))
            joint1 = yield from trial(p0, step)
        return step

    # ---- tree building ----------------------------------------------------------------------
    def leaf(self, x, p, log_slice, step, joint0):
        x1, p1 = yield from self.leapfrog(x, p, step)
        joint1 = (yield (LOGPDF, x1)) - 0.5 * np.inner(p1, p1)
        in_slice = float(log_slice <= joint1)
        ok = log_slice < (_MAX_ENERGY_ERROR + joint1)
        outside = False
        if ok:
            mh = min(1., np.exp(joint1 - joint0))
        else:
            mh = 0.
            if np.isinf((yield (LOGPDF, x1))):   # zero density: outside the support, not divergence
                outside = True
            else:
                logger.debug("NUTS: Diverging error: log_joint={}, params={}, params1={}, "
                             "momentum={}, momentum1={}.".format(joint1, x, x1, p, p1))
        return _Tree(x1, p1, x1, p1, x1, in_slice, ok, mh, 1., not ok, outside)

    def subtree(self, x, p, log_slice, step, depth, joint0):
        if depth == 0:
            return (yield from self.leaf(x, p, log_slice, step, joint0))
        first = yield from self.subtree(x, p, log_slice, step, depth - 1, joint0)
        if not first.ok:
            return first
        if step < 0:
            second = yield from self.subtree(first.left, first.p_left, log_slice, step, depth - 1,
                                             joint0)
            left, p_left, right, p_right = second.left, second.p_left, first.right, first.p_right
        else:
            second = yield from self.subtree(first.right, first.p_right, log_slice, step,
                                             depth - 1, joint0)
            left, p_left, right, p_right = first.left, first.p_left, second.right, second.p_right
        candidate = first.candidate
        if second.n_slice > 0:
            if float(second.n_slice) / (first.n_slice + second.n_slice) > self.rs.rand():
                candidate = second.candidate
        ok = second.ok and _no_u_turn(left, p_left, right, p_right)
        return _Tree(left, p_left, right, p_right, candidate, first.n_slice + second.n_slice, ok,
                     first.mh_sum + second.mh_sum, first.n_steps + second.n_steps,
                     second.diverged, second.outside)


def nuts_chain(n_iter, params0, n_adapt=None, target_prob=0.6, max_depth=5, seed=0, info_freq=100,
               max_retry_inits=20, stepsize=None):
    """Coroutine form of `nuts`: yields (kind, point) requests, returns the (n_iter, dim) chain."""
    params0 = np.asarray(params0, dtype=float)
    random_state = np.random.RandomState(seed)
    n_adapt = n_adapt if n_adapt is not None else n_iter // 2
    logger.info("NUTS: Performing {} iterations with {} adaptation steps.".format(n_iter, n_adapt))
    target0 = yield (LOGPDF, params0)
    if np.isinf(target0):
        raise ValueError("NUTS: Bad initialization point {}, logpdf -> -inf.".format(params0))
    sampler = _Nuts(random_state)
    if stepsize is None:
        stepsize = yield from sampler.initial_stepsize(params0, target0, max_retry_inits)
    logger.debug("NUTS: Set initial stepsize {}.".format(stepsize))

    # dual averaging (Hoffman & Gelman, section 3.2)
    mu = np.log(10. * stepsize)
    log_avg_stepsize = 0.
    h_bar = 0.
    gamma, t0, kappa = 0.05, 10., 0.75

    samples = np.empty((n_iter + 1,) + params0.shape)
    samples[0, :] = params0
    n_diverged = n_outside = n_total = 0
    for it in range(1, n_iter + 1):
        p0 = random_state.randn(*params0.shape)
        current = samples[it - 1, :]
        joint0 = (yield (LOGPDF, current)) - 0.5 * np.inner(p0, p0)
        log_slice = joint0 - random_state.exponential()
        samples[it, :] = current
        left = right = current
        p_left = p_right = p0
        n_slice = 1
        depth = 0
        keep_going = True
        while keep_going and depth <= max_depth:
            if random_state.rand() < 0.5:
                tree = yield from sampler.subtree(right, p_right, log_slice, stepsize, depth, joint0)
                right, p_right = tree.right, tree.p_right
            else:
                tree = yield from sampler.subtree(left, p_left, log_slice, -stepsize, depth, joint0)
                left, p_left = tree.left, tree.p_left
            if tree.ok == 1:
                if random_state.rand() < float(tree.n_slice) / n_slice:
                    samples[it, :] = tree.candidate
            n_slice += tree.n_slice
            if not tree.outside:
                n_diverged += tree.diverged
            n_outside += tree.outside
            n_total += tree.n_steps
            keep_going = tree.ok and _no_u_turn(left, p_left, right, p_right)
            depth += 1
            if depth > max_depth:
                logger.debug("NUTS: Maximum recursion depth {} exceeded.".format(max_depth))

        if it <= n_adapt:
            h_bar = (1. - 1. / (it + t0)) * h_bar \
                + (target_prob - float(tree.mh_sum) / tree.n_steps) / (it + t0)
            log_stepsize = mu - np.sqrt(it) / gamma * h_bar
            log_avg_stepsize = it ** (-kappa) * log_stepsize + \
                (1. - it ** (-kappa)) * log_avg_stepsize
            stepsize = np.exp(log_stepsize)
        elif it == n_adapt + 1:
            stepsize = np.exp(log_avg_stepsize)
            n_diverged = n_outside = n_total = 0
            logger.info("NUTS: Adaptation/warmup finished. Sampling...")
            logger.debug("NUTS: Set final stepsize {}.".format(stepsize))
        if it % info_freq == 0 and it < n_iter:
            logger.info("NUTS: Iterations performed: {}/{}...".format(it, n_iter))

    info = "NUTS: Acceptance ratio: {:.3f}".format(float(n_iter - n_adapt) / n_total)
    if n_outside > 0:
        info += ". After warmup {} proposals were outside of the region allowed by priors " \
                "and rejected, decreasing acceptance ratio.".format(n_outside)
    logger.info(info)
    if n_diverged > 0:
        logger.warning("NUTS: Diverged proposals after warmup (i.e. n_adapt={} steps): {}".format(
            n_adapt, n_diverged))
    return samples[1:, :]


def metropolis_chain(n_samples, params0, sigma_proposals, warmup=0, seed=0):
    """Coroutine form of `metropolis`."""
    params0 = np.asarray(params0, dtype=float)
    random_state = np.random.RandomState(seed)
    total = n_samples + warmup
    chain = np.empty((total + 1,) + params0.shape)
    chain[0, :] = params0
    logp = yield (LOGPDF, params0)
    if np.isinf(logp):
        raise ValueError(
            "Metropolis: Bad initialization point {},logpdf -> -inf.".format(params0))
    n_accepted = 0
    for it in range(1, total + 1):
        chain[it, :] = chain[it - 1, :] + sigma_proposals * random_state.randn(*params0.shape)
        logp_new = yield (LOGPDF, chain[it, :])
        rejected = (np.exp(logp_new - logp) < random_state.rand()) or np.isinf(logp_new) \
            or np.isnan(logp_new)
        if rejected:
            chain[it, :] = chain[it - 1, :]
        else:
            logp = logp_new
            n_accepted += 1
    logger.info("{}: Total acceptance ratio: {:.3f}".format(__name__, float(n_accepted) / total))
    return chain[(1 + warmup):, :]


# ----------------------------------------------------------------------------------- drivers
def _run_single(chain, target, grad_target):
    """Drive one chain coroutine with plain callables, one evaluation per request."""
    try:
        kind, x = next(chain)
        while True:
            kind, x = chain.send(target(x) if kind == LOGPDF else grad_target(x))
    except StopIteration as done:
        return done.value


def nuts(n_iter, params0, target, grad_target, n_adapt=None, target_prob=0.6, max_depth=5, seed=0,
         info_freq=100, max_retry_inits=20, stepsize=None):
    """Sample `target` (a log density) with NUTS; returns the (n_iter, dim) chain including the
    adaptation iterations (mcmc.py:105-299).

    n_adapt : dual-averaging iterations (default n_iter // 2); target_prob : desired mean
    acceptance (delta); max_depth : maximum number of doublings; stepsize : initial step size
    (found by trial and error when None)."""
    return _run_single(nuts_chain(n_iter, params0, n_adapt=n_adapt, target_prob=target_prob,
                                  max_depth=max_depth, seed=seed, info_freq=info_freq,
                                  max_retry_inits=max_retry_inits, stepsize=stepsize),
                       target, grad_target)


def metropolis(n_samples, params0, target, sigma_proposals, warmup=0, seed=0):
    """Random-walk Metropolis with Gaussian proposals of standard deviation `sigma_proposals`
    (mcmc.py:379-429); returns the (n_samples, dim) chain after `warmup` discarded iterations."""
    return _run_single(metropolis_chain(n_samples, params0, sigma_proposals, warmup=warmup,
                                        seed=seed), target, None)


class _PointCache:
    """The last few evaluated points of one chain: NUTS asks for the gradient and the density of
    the same point back to back, and a leaf starts where the previous one ended."""

    def __init__(self, size=4):
        self.entries, self.size = [], size

    def lookup(self, kind, x):
        """The cached log density / gradient of x, or None when it has to be evaluated."""
        key = np.asarray(x).tobytes()
        for k, logpdf, grad in self.entries:
            if k == key:
                return logpdf if kind == LOGPDF else grad
        return None

    def store(self, x, logpdf, grad):
        key = np.asarray(x).tobytes()
        # a point first seen in a density-only round is stored without a gradient: replace that
        # entry, or the later gradient request keeps missing and the point is re-evaluated
        self.entries = [e for e in self.entries if e[0] != key]
        self.entries.append((key, logpdf, grad))
        if len(self.entries) > self.size:
            del self.entries[0]


def run_lockstep(chains, evaluate):
    """Advance several chain coroutines together.

    evaluate(X (k, dim), with_grad) -> (logpdf (k,), grad (k, dim) or None) is called once per
    round with the points the chains are waiting for (those not answered from their caches), e.g.
    one batched device launch; gradients are requested whenever any waiting chain needs one.
    Returns the chain results in order."""
    chains = list(chains)
    results = [None] * len(chains)
    caches = [_PointCache() for _ in chains]
    pending = {i: next(chain) for i, chain in enumerate(chains)}
    while pending:
        need = [i for i, (kind, x) in pending.items() if caches[i].lookup(kind, x) is None]
        if need:
            with_grad = any(pending[i][0] == GRAD for i in need)
            X = np.array([np.asarray(pending[i][1], dtype=float) for i in need])
            logpdf, grad = evaluate(X, with_grad)
            for row, i in enumerate(need):
                caches[i].store(pending[i][1], logpdf[row],
                                np.array(grad[row]) if with_grad else None)
        for i in list(pending):
            kind, x = pending[i]
            while True:    # keep answering from the cache while the chain asks for known points
                value = caches[i].lookup(kind, x)
                if value is None:
                    pending[i] = (kind, x)
                    break
                try:
                    kind, x = chains[i].send(value)
                except StopIteration as done:
                    results[i] = done.value
                    del pending[i]
                    break
    return results
