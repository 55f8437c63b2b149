"""BOLFI on the device: GP surrogate, LCBSC acquisition, Bayesian optimisation loop.

Mirrors (file:line in /root/reference):
  GPyRegression          elfi/methods/bo/gpy_regression.py:15-364   (duck type used by acquisition
                         and posterior code: input_dim, bounds, parameter_names, X, Y, noise,
                         n_evidence, is_sampling, predict, predict_mean, predictive_gradients,
                         predictive_gradient_mean, update, optimize, copy)
  AcquisitionBase, LCBSC elfi/methods/bo/acquisition.py:16-301
  minimize               elfi/methods/bo/utils.py:40-111
  BayesianOptimization, BOLFI.fit/extract_posterior   elfi/methods/inference/bolfi.py:26-462
  BolfiPosterior (logpdf / pdf core)                  elfi/methods/posteriors.py:21-189

The GP arithmetic (Gram, Cholesky, inverse factor, predict, gradients, LCBSC) runs in gp.cu.
Kernel: RBF + Bias with the reference's default hyper-parameter heuristics
(gpy_regression.py:255-264).  Hyper-parameter optimisation: GPy (SCG on the marginal likelihood
with Gamma priors) is not available to pin against -> `optimize()` maximises the same objective
(log marginal likelihood + Gamma log-priors) with L-BFGS-B on log-parameters; PARITY UNPINNED.
"""
import copy
import logging
from math import ceil

import numpy as np
import scipy.optimize
import scipy.stats as ss
import torch

from . import _lib
from . import device as dev
from .results import OptimizationResult
from .samplers import ModelPrior, ParameterInference

logger = logging.getLogger(__name__)

JITTER = 1e-8   # GPy's exact Gaussian inference adds 1e-8 to the diagonal


class GPyRegression:
    """GP regression with an RBF + Bias kernel and Gaussian noise, fitted on the device.

    Named after the reference class it stands in for; no GPy involved."""

    def __init__(self, parameter_names=None, bounds=None, optimizer="lbfgsb", max_opt_iters=50,
                 gp=None, **gp_params):
        if parameter_names is None:
            input_dim = 1
        elif isinstance(parameter_names, (list, tuple)):
            input_dim = len(parameter_names)
        else:
            raise ValueError("Keyword `parameter_names` must be a list of strings")
        if bounds is None:
            logger.warning('Parameter bounds not specified. Using [0,1] for each parameter.')
            bounds = [(0, 1)] * input_dim
        elif len(bounds) != input_dim:
            raise ValueError('Length of `bounds` ({}) does not match the length of '
                             '`parameter_names` ({}).'.format(len(bounds), input_dim))
        elif isinstance(bounds, dict):
            if len(bounds) == 1:
                bounds = [bounds[n] for n in bounds.keys()]
            else:
                bounds = [bounds[n] for n in parameter_names]
        else:
            raise ValueError("Keyword `bounds` must be a dictionary "
                             "`{'parameter_name': (lower, upper), ... }`")
        self.parameter_names = parameter_names
        self.input_dim = input_dim
        self.bounds = bounds
        self.gp_params = gp_params
        self.optimizer = optimizer
        self.max_opt_iters = max_opt_iters
        self.is_sampling = False
        self._X = None          # host (n, p)
        self._Y = None          # host (n, 1)
        self._hyper = None      # dict(kernel_var, lengthscale, bias_var, noise_var)
        self._priors = None     # Gamma prior (a, b) per hyper-parameter
        self._factor = None     # device tensors of the current fit

    # ---- data / state ---------------------------------------------------------------------
    @property
    def n_evidence(self):
        return 0 if self._X is None else len(self._X)

    @property
    def X(self):
        return self._X

    @property
    def Y(self):
        return self._Y

    @property
    def noise(self):
        return self._hyper['noise_var']

    @property
    def hyperparameters(self):
        return dict(self._hyper)

    def _default_hyper(self, x, y):
        """gpy_regression.py:255-280: heuristics from the bounds and the first data."""
        length_scale = (np.max(self.bounds) - np.min(self.bounds)) / 3.
        kernel_var = (np.max(y) / 3.) ** 2.
        bias_var = kernel_var / 4.
        noise_var = self.gp_params.get('noise_var') or np.max(y) ** 2. / 100.
        self._hyper = dict(kernel_var=float(kernel_var), lengthscale=float(length_scale),
                           bias_var=float(bias_var), noise_var=float(noise_var))
        for k in ('kernel_var', 'lengthscale', 'bias_var'):
            if k in self.gp_params and self.gp_params[k] is not None:
                self._hyper[k] = float(self.gp_params[k])
        # Gamma.from_EV(E, V) with E = V = value -> shape a = E^2/V = E, rate b = E/V = 1
        self._priors = {'lengthscale': (length_scale, 1.0), 'kernel_var': (kernel_var, 1.0),
                        'bias_var': (bias_var, 1.0)}

    def update(self, x, y, optimize=False):
        """Append evidence and refit (the reference rebuilds the GP on every update, 286-315)."""
        x = np.asarray(dev.to_host(x), dtype=np.float64).reshape((-1, self.input_dim))
        y = np.asarray(dev.to_host(y), dtype=np.float64).reshape((-1, 1))
        if self._X is None:
            self._default_hyper(x, y)
            self._X, self._Y = x, y
        else:
            self._X = np.r_[self._X, x]
            self._Y = np.r_[self._Y, y]
        self._fit()
        if optimize:
            self.optimize()

    def _fit(self, hyper=None):
        h = hyper or self._hyper
        n, p = self._X.shape
        n_pad = int(_lib.load().elfi_b200_gp_padded_size(n))
        Xd = dev.to_device(self._X)
        yd = dev.to_device(self._Y.reshape(-1))
        f = self._factor
        if f is None or f['n_pad'] != n_pad:
            f = dict(n_pad=n_pad, L=dev.empty((n_pad, n_pad)), W=dev.empty((n_pad, n_pad)),
                     U=dev.empty((n_pad, n_pad)))
        f.update(X=Xd, y=yd, n=n, alpha=dev.empty((n,)),
                 info=dev.zeros((1,), dtype=torch.int32))
        _lib.call('elfi_b200_gp_fit_f64', dev.context(), dev.ptr(Xd), p, dev.ptr(yd), n, p,
                  h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'] + JITTER,
                  dev.ptr(f['L']), dev.ptr(f['W']), dev.ptr(f['U']), n_pad, dev.ptr(f['alpha']),
                  dev.ptr(f['info']), dev.stream_ptr())
        info = int(f['info'].item())
        if info != 0:
            raise np.linalg.LinAlgError('Cholesky failed: non-positive pivot at {}'.format(info - 1))
        f['hyper'] = dict(h)
        self._factor = f
        return f

    # ---- prediction -----------------------------------------------------------------------
    def predict(self, x, noiseless=False):
        """GP mean and variance at x -> ((m, 1), (m, 1)) host arrays (gpy_regression.py:98-149)."""
        x = np.asanyarray(dev.to_host(x), dtype=np.float64).reshape((-1, self.input_dim))
        if self._factor is None:
            return np.zeros((x.shape[0], 1)), np.ones((x.shape[0], 1))
        mean, var, _ = self.predict_device(x, noiseless=noiseless)
        return mean.cpu().numpy()[:, None], var.cpu().numpy()[:, None]

    def predict_device(self, x, noiseless=True, beta=None):
        """Device tensors (mean, var, acq): acq = LCBSC value when beta is given."""
        f, h = self._factor, self._factor['hyper']
        xq = dev.to_device(x).reshape(-1, self.input_dim)
        m = xq.shape[0]
        mean, var = dev.empty((m,)), dev.empty((m,))
        acq = dev.empty((m,)) if beta is not None else None
        _lib.call('elfi_b200_gp_predict_f64', dev.context(), dev.ptr(xq), self.input_dim, m,
                  dev.ptr(f['X']), self.input_dim, f['n'], self.input_dim, dev.ptr(f['W']),
                  f['n_pad'], dev.ptr(f['alpha']), h['kernel_var'], h['lengthscale'],
                  h['bias_var'], 0.0 if noiseless else h['noise_var'],
                  float(beta) if beta is not None else 0.0, dev.ptr(mean), dev.ptr(var),
                  dev.ptr(acq), dev.stream_ptr())
        return mean, var, acq

    def predict_mean(self, x):
        return self.predict(x)[0]

    def predictive_gradients(self, x):
        """Gradients of the GP mean and variance -> ((m, p), (m, p)) (gpy_regression.py:186-223)."""
        x = np.asanyarray(dev.to_host(x), dtype=np.float64).reshape((-1, self.input_dim))
        if self._factor is None:
            return np.zeros((x.shape[0], self.input_dim)), np.zeros((x.shape[0], self.input_dim))
        _, _, gm, gv = self._predict_grad_device(x)
        return gm.cpu().numpy(), gv.cpu().numpy()

    def _predict_grad_device(self, x):
        f, h = self._factor, self._factor['hyper']
        xq = dev.to_device(x).reshape(-1, self.input_dim)
        m, p = xq.shape
        mean, var = dev.empty((m,)), dev.empty((m,))
        gm, gv = dev.empty((m, p)), dev.empty((m, p))
        _lib.call('elfi_b200_gp_predict_grad_f64', dev.context(), dev.ptr(xq), p, m,
                  dev.ptr(f['X']), p, f['n'], p, dev.ptr(f['W']), dev.ptr(f['U']), f['n_pad'],
                  dev.ptr(f['alpha']), h['kernel_var'], h['lengthscale'], h['bias_var'],
                  dev.ptr(mean), dev.ptr(var), dev.ptr(gm), dev.ptr(gv), dev.stream_ptr())
        return mean, var, gm, gv

    def predictive_gradient_mean(self, x):
        return self.predictive_gradients(x)[0]

    # ---- hyper-parameters -------------------------------------------------------------------
    def log_marginal_likelihood(self, hyper=None):
        """-1/2 y^T alpha - sum log L_ii - n/2 log 2 pi for the given (or current) hyper-parameters."""
        f = self._fit(hyper) if hyper is not None else self._factor
        n = f['n']
        diag = torch.diagonal(f['L'])[:n]
        return float(-0.5 * torch.dot(f['y'], f['alpha']) - torch.log(diag).sum()
                     - 0.5 * n * np.log(2 * np.pi))

    def optimize(self):
        """Maximise log marginal likelihood + Gamma log-priors over (kernel_var, lengthscale,
        bias_var, noise_var) in log space.  PARITY UNPINNED w.r.t. GPy's SCG."""
        names = ['kernel_var', 'lengthscale', 'bias_var', 'noise_var']
        x0 = np.log([self._hyper[k] for k in names])

        def objective(logh):
            return -self.log_posterior_hyper(dict(zip(names, np.exp(logh))))
        f0 = objective(x0)
        res = scipy.optimize.minimize(objective, x0, method='L-BFGS-B',
                                      bounds=[(v - 12.0, v + 12.0) for v in x0],
                                      options={'maxiter': self.max_opt_iters})
        if np.isfinite(res.fun) and res.fun <= f0:
            self._hyper = dict(zip(names, np.exp(res.x).tolist()))
        else:
            logger.warning("Numerical error in GP optimization. Stopping optimization")
        self._fit()

    def log_posterior_hyper(self, hyper=None):
        """log marginal likelihood + Gamma log-priors (the quantity `optimize` maximises)."""
        h = hyper or self._hyper
        try:
            val = self.log_marginal_likelihood(dict(h))
        except (np.linalg.LinAlgError, _lib.ElfiB200Error):
            return -1e25
        for k, (a, b) in self._priors.items():
            val += ss.gamma.logpdf(h[k], a=a, scale=1.0 / b)
        return val if np.isfinite(val) else -1e25

    def copy(self):
        kopy = copy.copy(self)
        if self._factor is not None:
            kopy._factor = {k: (v.clone() if isinstance(v, torch.Tensor) else copy.copy(v))
                            for k, v in self._factor.items()}
        kopy._hyper = copy.copy(self._hyper)
        return kopy


# ------------------------------------------------------------------------------ acquisition
def minimize(fun, bounds, method='L-BFGS-B', constraints=None, grad=None, prior=None,
             n_start_points=10, maxiter=1000, random_state=None):
    """Multi-start local minimisation (elfi/methods/bo/utils.py:40-111)."""
    ndim = len(bounds)
    start_points = np.empty((n_start_points, ndim))
    if prior is None:
        random_state = random_state or np.random
        for i in range(ndim):
            start_points[:, i] = random_state.uniform(*bounds[i], n_start_points)
    else:
        start_points = prior.rvs(n_start_points, random_state=random_state)
        if len(start_points.shape) == 1:
            start_points = start_points[:, None]
        for i in range(ndim):
            start_points[:, i] = np.clip(start_points[:, i], *bounds[i])
    locs, vals = [], np.empty(n_start_points)
    for i in range(n_start_points):
        result = scipy.optimize.minimize(fun, start_points[i, :], method=method, jac=grad,
                                         bounds=bounds, constraints=constraints,
                                         options={'maxiter': maxiter})
        locs.append(result['x'])
        vals[i] = result['fun']
    ind_min = np.argmin(vals)
    locs_out = locs[ind_min]
    for i in range(ndim):
        locs_out[i] = np.clip(locs_out[i], *bounds[i])
    return locs[ind_min], vals[ind_min]


class AcquisitionBase:
    """elfi/methods/bo/acquisition.py:16-191."""

    def __init__(self, model, prior=None, n_inits=10, max_opt_iters=1000, noise_var=None,
                 exploration_rate=10, seed=None, constraints=None):
        self.model = model
        self.prior = prior
        self.n_inits = int(n_inits)
        self.max_opt_iters = int(max_opt_iters)
        self.constraints = constraints
        if noise_var is not None:
            self._check_noise_var(noise_var)
            if isinstance(noise_var, dict):
                noise_var = list(map(noise_var.get, self.model.parameter_names))
        self.noise_var = noise_var
        self.exploration_rate = exploration_rate
        self.random_state = np.random if seed is None else np.random.RandomState(seed)
        self.seed = 0 if seed is None else seed

    def _check_noise_var(self, noise_var):
        if isinstance(noise_var, dict):
            if not set(noise_var) == set(self.model.parameter_names):
                raise ValueError("Acquisition noise dictionary should contain all parameters.")
            if not all(isinstance(x, (int, float)) for x in noise_var.values()):
                raise ValueError("Acquisition noise dictionary values should all be int or float.")
            if any([x < 0 for x in noise_var.values()]):
                raise ValueError("Acquisition noises values should all be "
                                 "non-negative int or float.")
        elif isinstance(noise_var, (int, float)):
            if noise_var < 0:
                raise ValueError("Acquisition noise should be non-negative int or float.")
        else:
            raise ValueError("Either acquisition noise is a float or it is a dictionary of "
                             "floats defining variance for each parameter dimension.")

    def evaluate(self, x, t=None):
        raise NotImplementedError

    def evaluate_gradient(self, x, t=None):
        raise NotImplementedError

    def acquire(self, n, t=None):
        xhat, _ = minimize(lambda x: self.evaluate(x, t), self.model.bounds,
                           method='L-BFGS-B' if self.constraints is None else 'SLSQP',
                           constraints=self.constraints,
                           grad=lambda x: self.evaluate_gradient(x, t), prior=self.prior,
                           n_start_points=self.n_inits, maxiter=self.max_opt_iters,
                           random_state=self.random_state)
        x = np.tile(xhat, (n, 1))
        return self._add_noise(x)

    def _add_noise(self, x):
        if self.noise_var is not None:
            noise_var = np.asanyarray(self.noise_var)
            if noise_var.ndim == 0:
                noise_var = np.tile(noise_var, self.model.input_dim)
            for i in range(self.model.input_dim):
                std = np.sqrt(noise_var[i])
                if std == 0:
                    continue
                xi = x[:, i]
                a = (self.model.bounds[i][0] - xi) / std
                b = (self.model.bounds[i][1] - xi) / std
                x[:, i] = ss.truncnorm.rvs(a, b, loc=xi, scale=std, size=len(x),
                                           random_state=self.random_state)
        return x


class LCBSC(AcquisitionBase):
    """Lower confidence bound selection criterion (acquisition.py:194-301); the value and its
    gradient are evaluated on the device, fused with the GP prediction."""

    def __init__(self, *args, delta=None, additive_cost=None, **kwargs):
        if delta is not None:
            if delta <= 0 or delta >= 1:
                logger.warning('Parameter delta should be in the interval (0,1)')
            kwargs['exploration_rate'] = 1 / delta
        super().__init__(*args, **kwargs)
        self.name = 'lcbsc'
        self.label_fn = 'Confidence Bound'
        self.additive_cost = additive_cost

    @property
    def delta(self):
        return 1 / self.exploration_rate

    def _beta(self, t):
        t += 1
        d = self.model.input_dim
        return 2 * np.log(t ** (2 * d + 2) * np.pi ** 2 / (3 * self.delta))

    def evaluate(self, x, t=None):
        """(m, 1) host array; use evaluate_device for large grids."""
        value = self.evaluate_device(x, t).cpu().numpy()[:, None]
        if self.additive_cost is not None:
            value += self.additive_cost.evaluate(x)
        return value

    def evaluate_device(self, x, t=None):
        if self.model._factor is None:
            m = np.asarray(dev.to_host(x)).reshape(-1, self.model.input_dim).shape[0]
            return dev.full((m,), -np.sqrt(self._beta(t)))
        return self.model.predict_device(x, noiseless=True, beta=self._beta(t))[2]

    def evaluate_gradient(self, x, t=None):
        x = np.asanyarray(x, dtype=np.float64).reshape((-1, self.model.input_dim))
        if self.model._factor is None:
            return np.zeros_like(x)
        mean, var, gm, gv = self.model._predict_grad_device(x)
        m, p = gm.shape
        gacq = dev.empty((m, p))
        _lib.call('elfi_b200_lcbsc_f64', dev.context(), dev.ptr(mean), dev.ptr(var), dev.ptr(gm),
                  dev.ptr(gv), m, p, float(self._beta(t)), None, dev.ptr(gacq), dev.stream_ptr())
        value = gacq.cpu().numpy()
        if self.additive_cost is not None:
            value += self.additive_cost.evaluate_gradient(x)
        return value


# ------------------------------------------------------------------------------------ BO loop
def ceil_to_batch_size(num, batch_size):
    return int(batch_size * ceil(num / batch_size))


class BayesianOptimization(ParameterInference):
    """elfi/methods/inference/bolfi.py:26-292 (sequential acquisitions)."""

    def __init__(self, model, target_name=None, bounds=None, initial_evidence=None,
                 update_interval=10, target_model=None, acquisition_method=None, acq_noise_var=0,
                 exploration_rate=10, batch_size=1, batches_per_acquisition=None, async_acq=False,
                 **kwargs):
        model, target_name = self._resolve_model(model, target_name)
        output_names = [target_name] + model.parameter_names
        super().__init__(model, output_names, batch_size=batch_size, **kwargs)
        target_model = target_model or GPyRegression(self.model.parameter_names, bounds=bounds)
        self.target_name = target_name
        self.target_model = target_model
        n_precomputed = 0
        n_initial, precomputed = self._resolve_initial_evidence(initial_evidence)
        if precomputed is not None:
            params = np.column_stack([precomputed[n] for n in self.target_model.parameter_names])
            n_precomputed = len(params)
            self.target_model.update(params, precomputed[target_name])
        self.batches_per_acquisition = batches_per_acquisition or self.max_parallel_batches
        prior = ModelPrior(self.model, parameter_names=self.target_model.parameter_names)
        self.acquisition_method = acquisition_method or LCBSC(
            self.target_model, prior=prior, noise_var=acq_noise_var,
            exploration_rate=exploration_rate, seed=self.seed)
        self.n_initial_evidence = n_initial
        self.n_precomputed_evidence = n_precomputed
        self.update_interval = update_interval
        self.async_acq = async_acq
        self.state['n_evidence'] = self.n_precomputed_evidence
        self.state['last_GP_update'] = self.n_initial_evidence
        self.state['acquisition'] = []

    def _resolve_initial_evidence(self, initial_evidence):
        precomputed = None
        n_required = max(10, 2 ** self.target_model.input_dim + 1)
        n_required = ceil_to_batch_size(n_required, self.batch_size)
        if initial_evidence is None:
            n_initial_evidence = n_required
        elif np.isscalar(initial_evidence):
            n_initial_evidence = int(initial_evidence)
        else:
            precomputed = initial_evidence
            n_initial_evidence = len(precomputed[self.target_name])
        if n_initial_evidence < 0:
            raise ValueError('Number of initial evidence must be positive or zero '
                             '(was {})'.format(initial_evidence))
        if precomputed is None and (n_initial_evidence % self.batch_size != 0):
            n_initial_evidence = ceil_to_batch_size(n_initial_evidence, self.batch_size)
        return n_initial_evidence, precomputed

    @property
    def n_evidence(self):
        return self.state.get('n_evidence', 0)

    @property
    def acq_batch_size(self):
        return self.batch_size * self.batches_per_acquisition

    def set_objective(self, n_evidence=None):
        if n_evidence is None:
            n_evidence = self.objective.get('n_evidence', self.n_evidence)
        self.objective['n_evidence'] = n_evidence
        self.objective['n_sim'] = n_evidence - self.n_precomputed_evidence

    def extract_result(self):
        def fun_1d(x):
            return self.target_model.predict_mean(x).ravel()
        result = scipy.optimize.differential_evolution(
            func=fun_1d, bounds=self.target_model.bounds, maxiter=1000, polish=True,
            init='latinhypercube', seed=self.seed)
        names = self.target_model.parameter_names
        batch_min = {p: result.x.reshape((-1, len(names)))[:, i] for i, p in enumerate(names)}
        outputs = {p: self.target_model.X[:, i] for i, p in enumerate(names)}
        outputs[self.target_name] = self.target_model.Y
        return OptimizationResult(x_min=batch_min, outputs=outputs,
                                  **self._extract_result_kwargs())

    def update(self, batch, batch_index):
        super().update(batch, batch_index)
        self.state['n_evidence'] += self.batch_size
        params = np.column_stack([np.asarray(dev.to_host(batch[n]))
                                  for n in self.target_model.parameter_names])
        optimize = self._should_optimize()
        self.target_model.update(params, dev.to_host(batch[self.target_name]), optimize)
        if optimize:
            self.state['last_GP_update'] = self.target_model.n_evidence

    def prepare_new_batch(self, batch_index):
        t = self._get_acquisition_index(batch_index)
        if t < 0:
            return
        acquisition = self.state['acquisition']
        if len(acquisition) == 0:
            acquisition = self.acquisition_method.acquire(self.acq_batch_size, t=t)
        names = self.target_model.parameter_names
        acq = np.asarray(acquisition[:self.batch_size]).reshape((-1, len(names)))
        self.state['acquisition'] = acquisition[self.batch_size:]
        return {p: acq[:, i] for i, p in enumerate(names)}

    def _get_acquisition_index(self, batch_index):
        acq_batch_size = self.batch_size * self.batches_per_acquisition
        initial_offset = self.n_initial_evidence - self.n_precomputed_evidence
        starting_sim_index = self.batch_size * batch_index
        return (starting_sim_index - initial_offset) // acq_batch_size

    def _should_optimize(self):
        current = self.target_model.n_evidence + self.batch_size
        next_update = self.state['last_GP_update'] + self.update_interval
        return current >= self.n_initial_evidence and current >= next_update


class BolfiPosterior:
    """Unnormalised BOLFI posterior  prior(x) * Phi((h - mu(x)) / sigma(x))
    (elfi/methods/posteriors.py:21-189; logpdf / pdf / unnormalised likelihood only)."""

    def __init__(self, model, threshold=None, prior=None, n_inits=10, max_opt_iters=1000, seed=0):
        self.model = model
        self.threshold = threshold
        self.prior = prior
        self.dim = self.model.input_dim
        if self.threshold is None:
            def fun_1d(x):
                return self.model.predict_mean(x).ravel()
            res = scipy.optimize.differential_evolution(func=fun_1d, bounds=self.model.bounds,
                                                        maxiter=1000, polish=True,
                                                        init='latinhypercube', seed=seed)
            self.threshold = float(res.fun)

    def _unnormalized_loglikelihood(self, x):
        x = np.asanyarray(x)
        ndim = x.ndim
        x = x.reshape((-1, self.dim))
        mean, var = self.model.predict(x, noiseless=True)
        logpdf = ss.norm.logcdf(self.threshold, mean, np.sqrt(var)).squeeze()
        if ndim == 0 or (ndim == 1 and self.dim > 1):
            logpdf = logpdf[0] if np.ndim(logpdf) else logpdf
        return logpdf

    def logpdf(self, x):
        return self._unnormalized_loglikelihood(x) + self.prior.logpdf(x)

    def pdf(self, x):
        return np.exp(self.logpdf(x))


class BOLFI(BayesianOptimization):
    """Bayesian optimisation for likelihood-free inference (bolfi.py:400-462)."""

    def fit(self, n_evidence, threshold=None, bar=True):
        if n_evidence is None:
            raise ValueError('You must specify the number of evidence (n_evidence) for the fitting')
        self.infer(n_evidence, bar=bar)
        return self.extract_posterior(threshold)

    def extract_posterior(self, threshold=None):
        if self.state['n_evidence'] == 0:
            raise ValueError('Model is not fitted yet, please see the `fit` method.')
        prior = ModelPrior(self.model, parameter_names=self.target_model.parameter_names)
        return BolfiPosterior(self.target_model, threshold=threshold, prior=prior)
