"""Result containers (mirror of elfi/methods/results.py:18-108, 387-413; plotting/saving omitted)."""
from collections import OrderedDict

import numpy as np


class ParameterInferenceResult:
    def __init__(self, method_name, outputs, parameter_names, **kwargs):
        self.method_name = method_name
        self.outputs = outputs.copy()
        self.parameter_names = parameter_names
        self.meta = kwargs

    @property
    def is_multivariate(self):
        for p in self.parameter_names:
            if self.outputs[p].ndim > 1:
                return True
        return False


class OptimizationResult(ParameterInferenceResult):
    def __init__(self, x_min, **kwargs):
        super().__init__(**kwargs)
        self.x_min = x_min


class Sample(ParameterInferenceResult):
    """Sampling result: `outputs` (dict of host arrays), `samples` (parameters only),
    `weights`, and meta available as attributes (threshold, n_sim, accept_rate, ...)."""

    def __init__(self, method_name, outputs, parameter_names, discrepancy_name=None, weights=None,
                 **kwargs):
        super().__init__(method_name=method_name, outputs=outputs,
                         parameter_names=parameter_names, **kwargs)
        self.samples = OrderedDict()
        for n in self.parameter_names:
            self.samples[n] = self.outputs[n]
        self.discrepancy_name = discrepancy_name
        self.weights = weights

    def __getattr__(self, item):
        if 'meta' in self.__dict__ and item in self.__dict__['meta']:
            return self.meta[item]
        raise AttributeError("No attribute '{}' in this sample".format(item))

    @property
    def n_samples(self):
        return len(self.outputs[self.parameter_names[0]])

    @property
    def dim(self):
        return len(self.parameter_names)

    @property
    def discrepancies(self):
        return None if self.discrepancy_name is None else self.outputs[self.discrepancy_name]

    @property
    def samples_array(self):
        return np.column_stack(tuple(self.samples.values()))

    @property
    def sample_means(self):
        return OrderedDict([(k, np.average(v, axis=0, weights=self.weights))
                            for k, v in self.samples.items()])

    @property
    def sample_means_array(self):
        return np.array(list(self.sample_means.values()))

    def __repr__(self):
        return 'Sample(method={}, n_samples={}, parameters={})'.format(
            self.method_name, self.n_samples, self.parameter_names)


class SmcSample(Sample):
    def __init__(self, method_name, outputs, parameter_names, populations, *args, **kwargs):
        super().__init__(method_name=method_name, outputs=outputs,
                         parameter_names=parameter_names, *args, **kwargs)
        self.populations = populations
        if self.weights is None:
            raise ValueError("No weights provided for the sample")

    @property
    def n_populations(self):
        return len(self.populations)
