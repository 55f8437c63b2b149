"""Result containers with the attribute surface of elfi/methods/results.py (Sample at 73-108,
SmcSample at 387-413, OptimizationResult at 55-70): host arrays in `outputs`, parameter draws in
`samples`, optional `weights`, and every extra keyword (threshold, n_sim, accept_rate, seed, cov,
means, adaptive_distance_w, ...) readable as an attribute.  Plotting / saving are out of scope.
"""
from collections import OrderedDict

import numpy as np


class _Result:
    """Common part: method name, node outputs, parameter names and free-form meta data."""

    def __init__(self, method_name, outputs, parameter_names, **meta):
        self.method_name = method_name
        self.outputs = dict(outputs)
        self.parameter_names = parameter_names
        self.meta = meta

    def __getattr__(self, name):
        # only reached when normal lookup fails: expose the meta entries as attributes
        meta = self.__dict__.get('meta')
        if meta is not None and name in meta:
            return meta[name]
        raise AttributeError("No attribute '{}' in this sample".format(name))

    @property
    def is_multivariate(self):
        return any(np.ndim(self.outputs[p]) > 1 for p in self.parameter_names)


ParameterInferenceResult = _Result


class OptimizationResult(_Result):
    """Result of an optimisation: `x_min` holds the minimiser per parameter."""

    def __init__(self, x_min, **kwargs):
        super().__init__(**kwargs)
        self.x_min = x_min


class Sample(_Result):
    """Draws from an (approximate) posterior."""

    def __init__(self, method_name, outputs, parameter_names, discrepancy_name=None, weights=None,
                 **meta):
        super().__init__(method_name, outputs, parameter_names, **meta)
        self.samples = OrderedDict((name, self.outputs[name]) for name in parameter_names)
        self.discrepancy_name = discrepancy_name
        self.weights = weights

    @property
    def n_samples(self):
        first = self.parameter_names[0]
        return len(self.outputs[first])

    @property
    def dim(self):
        return len(self.parameter_names)

    @property
    def discrepancies(self):
        if self.discrepancy_name is None:
            return None
        return self.outputs[self.discrepancy_name]

    @property
    def samples_array(self):
        return np.column_stack([self.samples[name] for name in self.parameter_names])

    @property
    def sample_means(self):
        means = OrderedDict()
        for name, draws in self.samples.items():
            means[name] = np.average(draws, axis=0, weights=self.weights)
        return means

    @property
    def sample_means_array(self):
        return np.array([v for v in self.sample_means.values()])

    def __repr__(self):
        return '{}(method={!r}, n_samples={}, parameters={})'.format(
            type(self).__name__, self.method_name, self.n_samples, list(self.parameter_names))


class SmcSample(Sample):
    """Final population of an SMC run plus the list of all populations."""

    def __init__(self, method_name, outputs, parameter_names, populations, *args, **kwargs):
        super().__init__(method_name, outputs, parameter_names, *args, **kwargs)
        if self.weights is None:
            raise ValueError("No weights provided for the sample")
        self.populations = populations

    @property
    def n_populations(self):
        return len(self.populations)


class BolfiSample(Sample):
    """Posterior draws of BOLFI.sample: `chains` is (n_chains, n_samples, n_parameters) with the
    warm-up iterations included; `samples` holds the post-warm-up draws of all chains
    (elfi/methods/results.py:507-543)."""

    def __init__(self, method_name, chains, parameter_names, warmup, **meta):
        chains = np.array(chains, copy=True)
        kept = chains[:, warmup:, :].reshape((-1,) + chains.shape[2:])
        outputs = {name: kept[:, i] for i, name in enumerate(parameter_names)}
        super().__init__(method_name=method_name, outputs=outputs, parameter_names=parameter_names,
                         chains=chains, n_chains=chains.shape[0], warmup=warmup, **meta)
