"""Host-side mirror of ELFI's operator (node) API for the hot path.

Same names, argument meaning and error behaviour as elfi/model/elfi_model.py, but only what
the path needs: a DAG of node states, the reference's deterministic execution order (so that
one per-batch ``RandomState`` is consumed exactly as in elfi/executor.py), and per-batch
execution in which the Summary / Distance operations run on the device.

What is mirrored (file:line in the reference):
  ElfiModel, generate            elfi/model/elfi_model.py:211-299
  NodeReference and node classes elfi/model/elfi_model.py:477-1151
  compile (observed twins, batch_size / meta / random_state feeders, pruning)
                                 elfi/compiler.py:32-235
  load (observed data, sub-seeded RandomState)   elfi/loader.py:31-178
  execution order + node calls   elfi/executor.py:44-246
  get_sub_seed                   elfi/utils.py:71-127
What is not: pools / stores, pickling, name inspection of the caller's frame, other clients.
"""
import uuid
from functools import partial

import networkx as nx
import numpy as np
import scipy.stats as ss

from . import device as dev
from . import ops

_default_model = None

SCIPY_ALIASES = {'normal': 'norm', 'exponential': 'expon', 'unif': 'uniform', 'bin': 'binom',
                 'binomial': 'binom'}


def get_default_model():
    global _default_model
    if _default_model is None:
        _default_model = ElfiModel()
    return _default_model


def set_default_model(model=None):
    global _default_model
    if model is None:
        model = ElfiModel()
    if not isinstance(model, ElfiModel):
        raise ValueError('{} is not an instance of ElfiModel'.format(ElfiModel))
    _default_model = model


def new_model(name=None, set_default=True):
    model = ElfiModel(name=name)
    if set_default:
        set_default_model(model)
    return model


def random_name(length=4, prefix=''):
    return prefix + str(uuid.uuid4().hex[0:length])


def observed_name(name):
    return "_{}_observed".format(name)


def is_observed_name(name):
    return isinstance(name, str) and name.startswith('_') and name.endswith('_observed')


def is_array(output):
    return hasattr(output, 'shape') and output.ndim > 0


def scipy_from_str(name):
    name = name.lower()
    name = SCIPY_ALIASES.get(name, name)
    return getattr(ss, name)


def get_sub_seed(seed, sub_seed_index, high=2 ** 31, cache=None):
    """Unique sub seed number `sub_seed_index` of `seed` (elfi/utils.py:71-127).

    Draws uint32 values below `high` from RandomState(seed) until sub_seed_index + 1 distinct
    ones have been seen; the last draw is the sub seed.  Bit-identical to the reference
    (golden values in tests/golden/meta.json)."""
    if isinstance(seed, np.random.RandomState):
        raise ValueError('Seed cannot be a random state')
    if sub_seed_index >= high:
        raise ValueError("Sub seed index {} is out of range".format(sub_seed_index))
    if cache and len(cache['seen']) < sub_seed_index + 1:
        random_state, seen = cache['random_state'], cache['seen']
    else:
        random_state, seen = np.random.RandomState(seed), set()
    draws = None
    wanted = sub_seed_index + 1
    while len(seen) != wanted:
        draws = random_state.randint(high, size=wanted - len(seen), dtype='uint32')
        seen.update(draws)
    if cache is not None:
        cache['random_state'] = random_state
        cache['seen'] = seen
    return draws[-1]


class ComputationContext:
    """batch_size + seed + optional output pool (+ caches) shared by all batches of one inference
    (elfi/model/elfi_model.py:126-208)."""

    def __init__(self, batch_size=None, seed=None, pool=None):
        if pool is not None and pool.has_context:
            if batch_size is None:
                batch_size = pool.batch_size
            elif batch_size != pool.batch_size:
                raise ValueError('Pool batch_size differs from the given batch_size!')
            if seed is None:
                seed = pool.seed
            elif seed != pool.seed:
                raise ValueError('Pool seed differs from the given seed!')
        self.batch_size = batch_size or 1
        self.seed = np.random.RandomState().get_state()[1][1] if seed is None else seed
        self.pool = pool
        self.caches = {'plan': {}, 'sub_seed': {}}
        self.num_submissions = 0
        if pool is not None and not pool.has_context:
            pool.set_context(self)

    def callback(self, batch, batch_index):
        """Store the finished batch in the pool (elfi_model.py:196-208)."""
        if self.pool is not None:
            self.pool.add_batch(batch, batch_index)


# ------------------------------------------------------------------------------------ graph
class ElfiModel:
    """A DAG of node states (mirror of elfi.ElfiModel for the hot path)."""

    def __init__(self, name=None, observed=None, source_net=None):
        self.source_net = source_net if source_net is not None else nx.DiGraph()
        self.source_net.graph.setdefault('name', name or "model_{}".format(random_name()))
        self.source_net.graph.setdefault('observed', observed or {})
        if name:
            self.source_net.graph['name'] = name

    # -- bookkeeping ---------------------------------------------------------------------
    @property
    def name(self):
        return self.source_net.graph['name']

    @name.setter
    def name(self, name):
        self.source_net.graph['name'] = name

    @property
    def observed(self):
        return self.source_net.graph['observed']

    @observed.setter
    def observed(self, observed):
        if not isinstance(observed, dict):
            raise ValueError("Observed data must be given in a dictionary with the node"
                             "name as the key")
        self.source_net.graph['observed'] = observed

    @property
    def nodes(self):
        return self.source_net.nodes()

    def has_node(self, name):
        return self.source_net.has_node(name)

    def add_node(self, name, state):
        if self.has_node(name):
            raise ValueError('Node {} already exists'.format(name))
        self.source_net.add_node(name, attr_dict=state)

    def get_node(self, name):
        return self.source_net.nodes[name]

    def get_state(self, name):
        return self.source_net.nodes[name]

    def get_parents(self, child_name):
        args = []
        for parent_name in self.source_net.predecessors(child_name):
            param = self.source_net[parent_name][child_name]['param']
            if isinstance(param, int):
                args.append((param, parent_name))
        return [a[1] for a in sorted(args)]

    def add_edge(self, parent_name, child_name, param_name=None):
        if param_name is None:
            param_name = len(self.get_parents(child_name))
        if not self.has_node(parent_name):
            raise ValueError('Parent {} does not exist'.format(parent_name))
        if not self.has_node(child_name):
            raise ValueError('Child {} does not exist'.format(child_name))
        self.source_net.add_edge(parent_name, child_name, param=param_name)

    def remove_node(self, name):
        if name in self.observed:
            self.observed.pop(name)
        parent_names = self.get_parents(name)
        self.source_net.remove_node(name)
        for p in parent_names:
            if p[0] == '_' and self.source_net.degree(p) == 0:
                self.remove_node(p)

    def update_node(self, name, updating_name):
        """`name` takes the state and parents of `updating_name` (NodeReference.become)."""
        obs = self.observed.pop(updating_name, None)
        out_edges = list(self.source_net.edges(name, data=True))
        self.remove_node(name)
        self.source_net.add_node(name, attr_dict=self.source_net.nodes[updating_name]['attr_dict'])
        self.source_net.add_edges_from(out_edges)
        for u, v, data in list(self.source_net.in_edges(updating_name, data=True)):
            self.source_net.add_edge(u, name, **data)
        self.remove_node(updating_name)
        if obs is not None:
            self.observed[name] = obs

    @property
    def parameter_names(self):
        return sorted([n for n in self.nodes if '_parameter' in self.get_state(n)['attr_dict']])

    @parameter_names.setter
    def parameter_names(self, parameter_names):
        parameter_names = set(parameter_names)
        for n in self.nodes:
            state = self.get_state(n)['attr_dict']
            if n in parameter_names:
                parameter_names.remove(n)
                state['_parameter'] = True
            else:
                state.pop('_parameter', None)
        if len(parameter_names) > 0:
            raise ValueError('Parameters {} not found from the model'.format(parameter_names))

    def copy(self):
        kopy = ElfiModel(source_net=nx.DiGraph(self.source_net))
        kopy.source_net.graph['observed'] = dict(self.observed)
        kopy.name = "{}_copy_{}".format(self.name, random_name())
        return kopy

    def get_reference(self, name):
        cls = self.get_node(name)['attr_dict']['_class']
        return cls.reference(name, self)

    def __getitem__(self, node_name):
        return self.get_reference(node_name)

    # -- execution -----------------------------------------------------------------------
    def generate(self, batch_size=1, outputs=None, with_values=None, seed=None):
        """Generate one batch of outputs (elfi/model/elfi_model.py:265-299)."""
        if outputs is None:
            outputs = list(self.source_net.nodes())
        elif isinstance(outputs, str):
            outputs = [outputs]
        if not isinstance(outputs, list):
            raise ValueError('Outputs must be a list of node names')
        if seed is None:
            seed = 'global'
        context = ComputationContext(batch_size, seed=seed)
        return execute_batch(self, outputs, context, 0, with_values)


# ------------------------------------------------------------------------- compile / execute
def _constant_topological_order(G):
    """Deterministic topological order: depth-first from the alphabetically sorted nodes,
    successors explored from the alphabetically last to the first, reverse post-order.
    Yields the same order as elfi/executor.py:162-246 for the same graph."""
    explored, post = set(), []
    for root in sorted(G.nodes()):
        if root in explored:
            continue
        stack = [(root, iter(sorted(G[root], reverse=True)))]
        on_path = {root}
        while stack:
            node, children = stack[-1]
            advanced = False
            for child in children:
                if child in explored:
                    continue
                if child in on_path:
                    raise nx.NetworkXUnfeasible("Graph contains a cycle.")
                on_path.add(child)
                stack.append((child, iter(sorted(G[child], reverse=True))))
                advanced = True
                break
            if not advanced:
                stack.pop()
                on_path.discard(node)
                explored.add(node)
                post.append(node)
    return post[::-1]


def compile_net(source_net, outputs):
    """source_net -> computation net (elfi/compiler.py): operations/outputs, observed twins,
    `_batch_size` / `_meta` / `_random_state` feeder nodes, pruned to the outputs' ancestors."""
    outputs = set(outputs)
    net = nx.DiGraph(outputs=outputs, name=source_net.graph['name'])
    net.add_nodes_from(source_net.nodes())
    net.add_edges_from(source_net.edges(data=True))
    for name, data in net.nodes(data=True):
        state = source_net.nodes[name]['attr_dict']
        if '_output' in state and '_operation' in state:
            raise ValueError("Cannot compile: both _output and _operation present "
                             "for node '{}'".format(name))
        if '_output' in state:
            data['output'] = state['_output']
        elif '_operation' in state:
            data['operation'] = state['_operation']
        else:
            raise ValueError("Cannot compile, no _output or _operation present for "
                             "node '{}'".format(name))
        if state.get('_uses_accept'):
            data['uses_accept'] = True

    observable, uses_observed = [], []
    for node in nx.topological_sort(source_net):
        state = source_net.nodes[node]['attr_dict']
        if state.get('_observable'):
            observable.append(node)
            net.add_node(observed_name(node), **{k: v for k, v in net.nodes[node].items()
                                                  if k != 'uses_accept'})
        elif state.get('_uses_observed'):
            uses_observed.append(node)
            net.add_node(observed_name(node), operation=lambda *a: tuple(a))
            net.add_edge(observed_name(node), node, param='observed')
        else:
            continue
        if not state.get('_stochastic'):
            for parent in source_net.predecessors(node):
                link = observed_name(parent) if parent in observable else parent
                net.add_edge(link, observed_name(node), **source_net[parent][node].copy())
    for node in uses_observed:
        for anc in nx.ancestors(net, observed_name(node)):
            if '_stochastic' in source_net.nodes.get(anc, {}).get('attr_dict', {}):
                raise ValueError("Observed nodes must be deterministic. Observed data depends "
                                 "on a non-deterministic node {}.".format(anc))

    for flag, feeder in (('_uses_batch_size', '_batch_size'), ('_uses_meta', '_meta')):
        for node, d in source_net.nodes(data=True):
            if d['attr_dict'].get(flag):
                if not net.has_node(feeder):
                    net.add_node(feeder)
                net.add_edge(feeder, node, param=feeder[1:])
    for node, d in source_net.nodes(data=True):
        if '_stochastic' in d['attr_dict']:
            if not net.has_node('_random_state'):
                net.add_node('_random_state')
            net.add_edge('_random_state', node, param='random_state')

    keep = set(outputs)
    for o in outputs:
        if not net.has_node(o):
            raise ValueError('Node {} is not in the model'.format(o))
        keep |= nx.ancestors(net, o)
    for node in list(net.nodes()):
        if node not in keep:
            net.remove_node(node)
    net.graph['order'] = _constant_topological_order(net)
    return net


def _call_node(net, node, values, accept):
    attr = net.nodes[node]
    op = attr['operation']
    args, kwargs = [], {}
    for parent in net.predecessors(node):
        param = net[parent][node]['param']
        if isinstance(param, int):
            args.append((param, values[parent]))
        else:
            kwargs[param] = values[parent]
    args = [a[1] for a in sorted(args, key=lambda t: t[0])]
    if accept is not None and attr.get('uses_accept') and node in accept:
        kwargs['accept'] = accept[node]
    try:
        return op(*args, **kwargs)
    except Exception as exc:
        raise exc.__class__("In executing node '{}': {}.".format(node, exc)).with_traceback(
            exc.__traceback__)


def execute_batch(model, outputs, context, batch_index, with_values=None, accept=None,
                  compiled=None):
    """compile (cached by the caller) + load + execute one batch; returns {name: output}.

    `accept` = {discrepancy_name: thresholds} asks a device Distance node to also return the
    accepted row indices (fused in the distance kernel); they come back under the key
    ('accepted', name)."""
    net = compiled if compiled is not None else compile_net(model.source_net, outputs)
    values = {}
    # ---- load (elfi/loader.py)
    observed = model.observed
    for name, obs in observed.items():
        if net.has_node(observed_name(name)):
            values[observed_name(name)] = obs
    if net.has_node('_batch_size'):
        values['_batch_size'] = context.batch_size
    if net.has_node('_meta'):
        values['_meta'] = {'batch_index': batch_index,
                           'submission_index': context.num_submissions,
                           'master_seed': context.seed, 'model_name': net.graph['name']}
    if net.has_node('_random_state'):
        seed = context.seed
        if isinstance(seed, str) and seed == 'global':
            values['_random_state'] = np.random.mtrand._rand
        elif isinstance(seed, (int, np.integer)):
            sub_seed = get_sub_seed(int(seed), batch_index, cache=context.caches.get('sub_seed'))
            values['_random_state'] = np.random.RandomState(sub_seed)
        else:
            raise ValueError("Seed of type {} is not supported".format(seed))
    # pool (elfi/loader.py:95-129): stored outputs replace their nodes, missing ones are
    # requested so that the callback can store them when the batch is done
    outputs = set(net.graph['outputs'])
    if context.pool is not None:
        stored = context.pool.get_batch(batch_index)
        for node in context.pool.stores:
            if not net.has_node(node):
                continue
            if node in stored:
                values[node] = stored[node]
            else:
                outputs.add(node)
    for k, v in (with_values or {}).items():
        if net.has_node(k):
            values[k] = v
    for node, attr in net.nodes(data=True):
        if 'output' in attr and node not in values:
            values[node] = attr['output']
    # observed twins are deterministic functions of the observed data (checked at compile time):
    # computed in the first batch of an inference, reused by the following ones (the reference
    # recomputes them per batch; here that would be kernel launches and a D2H per batch)
    obs_cache = context.caches.setdefault('observed', {}).setdefault(id(net), {})
    for node, out in obs_cache.items():
        values.setdefault(node, out)

    # ---- which nodes must run: ancestors of the outputs not cut off by a known value
    needed = [o for o in outputs if o not in values]
    todo = set()
    stack = list(needed)
    while stack:
        n = stack.pop()
        if n in todo or n in values:
            continue
        todo.add(n)
        stack.extend(net.predecessors(n))
    extras = {}
    for node in net.graph['order']:
        if node not in todo:
            continue
        if 'operation' not in net.nodes[node]:
            raise ValueError('Generative graph has no op or output present for node '
                             '{}'.format(node))
        out = _call_node(net, node, values, accept)
        if isinstance(out, AcceptedOutput):
            extras[('accepted', node)] = out.accepted
            out = out.value
        values[node] = out
        if is_observed_name(node) and compiled is not None:
            obs_cache[node] = out
    result = {k: values[k] for k in outputs}
    result.update(extras)
    return result


class AcceptedOutput:
    """Distance output + accepted row indices (fused acceptance)."""

    def __init__(self, value, accepted):
        self.value = value
        self.accepted = accepted


# ------------------------------------------------------------------------------------ nodes
class NodeReference:
    """Base class of node objects: a named handle on a state dict stored in the model
    (elfi/model/elfi_model.py:477-731)."""

    def __init__(self, *parents, state=None, model=None, name=None):
        state = state or {}
        state['_class'] = self.__class__
        model = self._determine_model(model, parents)
        name = self._give_name(name, model)
        model.add_node(name, state)
        self._init_reference(name, model)
        self._add_parents(parents)

    def _add_parents(self, parents):
        for parent in parents:
            if not isinstance(parent, NodeReference):
                parent_name = self._new_name('_' + self.name)
                parent = Constant(parent, name=parent_name, model=self.model)
            self.model.add_edge(parent.name, self.name)

    def _determine_model(self, model, parents):
        if not isinstance(model, ElfiModel) and model is not None:
            raise ValueError('Invalid model passed {}'.format(model))
        for p in parents:
            if isinstance(p, NodeReference):
                if model is None:
                    model = p.model
                elif model != p.model:
                    raise ValueError('Parents are from different models!')
        if model is None:
            model = get_default_model()
        return model

    @property
    def parents(self):
        return [self.model[p] for p in self.model.get_parents(self.name)]

    @classmethod
    def reference(cls, name, model):
        instance = cls.__new__(cls)
        instance._init_reference(name, model)
        return instance

    def become(self, other_node):
        if other_node.model is not self.model:
            raise ValueError('The other node belongs to a different model')
        self.model.update_node(self.name, other_node.name)
        _class = self.state['attr_dict'].get('_class', NodeReference)
        if not isinstance(self, _class):
            self.__class__ = _class
        other_node.name = self.name
        other_node.model = self.model

    def _init_reference(self, name, model):
        self.name = name
        self.model = model

    def generate(self, batch_size=1, with_values=None):
        result = self.model.generate(batch_size, self.name, with_values=with_values)
        return result[self.name]

    def _give_name(self, name, model):
        if name is not None:
            if name[-1] == '*':
                name = self._new_name(name[:-1], model)
            return name
        # the reference inspects the caller's source line for `x = elfi.Node(...)`; here an
        # explicit name is expected and a random one is generated otherwise
        return self._new_name(model=model)

    def _new_name(self, basename='', model=None):
        model = model or self.model
        if not basename:
            basename = '_{}'.format(self.__class__.__name__.lower())
        while True:
            name = "{}_{}".format(basename, random_name())
            if not model.has_node(name):
                break
        return name

    @property
    def state(self):
        if self.model is None:
            raise ValueError('{} {} is not initialized'.format(self.__class__.__name__, self.name))
        return self.model.get_node(self.name)

    def __getitem__(self, item):
        return self.state[item]

    def __setitem__(self, item, value):
        self.state[item] = value

    @property
    def uses_meta(self):
        return self.state['attr_dict'].get('_uses_meta', False)

    @uses_meta.setter
    def uses_meta(self, val):
        self.state['attr_dict']['_uses_meta'] = val

    def __repr__(self):
        return "{}(name='{}')".format(self.__class__.__name__, self.name)

    def __str__(self):
        return self.name


class Constant(NodeReference):
    def __init__(self, value, **kwargs):
        super().__init__(state=dict(_output=value), **kwargs)


class _HostOperation:
    """User-supplied Operation callables expect NumPy: device inputs are copied to the host
    first (explicit D2H; generic Operations are host logic, not part of the CUDA hot path)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *args, **kwargs):
        args = [dev.to_host(a) if dev.is_device_array(a) else a for a in args]
        kwargs = {k: (dev.to_host(v) if dev.is_device_array(v) else v) for k, v in kwargs.items()}
        return self.fn(*args, **kwargs)


class Operation(NodeReference):
    def __init__(self, fn, *parents, **kwargs):
        super().__init__(*parents, state=dict(_operation=_HostOperation(fn)), **kwargs)


def rvs_from_distribution(*params, batch_size, distribution, size=None, random_state=None):
    """elfi/model/utils.py:6-34."""
    size = (batch_size,) if size is None else (batch_size,) + size
    return distribution.rvs(*params, size=size, random_state=random_state)


class RandomVariable(NodeReference):
    def __init__(self, distribution, *params, size=None, **kwargs):
        state = dict(distribution=distribution, size=size, _uses_batch_size=True,
                     _stochastic=True)
        if not (size is None or isinstance(size, tuple)):
            size = (size,)
        dist = scipy_from_str(distribution) if isinstance(distribution, str) else distribution
        if not hasattr(dist, 'rvs'):
            raise ValueError("Distribution {} must implement a rvs method".format(distribution))
        state['_operation'] = partial(rvs_from_distribution, distribution=dist, size=size)
        super().__init__(*params, state=state, **kwargs)

    @property
    def distribution(self):
        distribution = self.state['attr_dict']['distribution']
        if isinstance(distribution, str):
            distribution = scipy_from_str(distribution)
        return distribution

    @property
    def size(self):
        return self.state['attr_dict']['size']


class Prior(RandomVariable):
    def __init__(self, distribution, *params, size=None, **kwargs):
        super().__init__(distribution, *params, size=size, **kwargs)
        self.state['attr_dict']['_parameter'] = True


class _Observable(NodeReference):
    def _set_observed(self, observed):
        if observed is not None:
            self.model.observed[self.name] = observed

    @property
    def observed(self):
        obs_name = observed_name(self.name)
        return self.model.generate(0, obs_name)[obs_name]


class Simulator(_Observable):
    """fn(*params, batch_size, random_state) -> array of length batch_size."""

    def __init__(self, fn, *params, observed=None, **kwargs):
        state = dict(_operation=fn, _uses_batch_size=True, _stochastic=True, _observable=True)
        super().__init__(*params, state=state, **kwargs)
        self._set_observed(observed)


class Summary(_Observable):
    """fn(*parents) -> summary statistic; may return a device array."""

    def __init__(self, fn, *parents, observed=None, **kwargs):
        if not parents:
            raise ValueError('This node requires that at least one parent is specified.')
        state = dict(_operation=fn, _observable=True)
        super().__init__(*parents, state=state, **kwargs)
        self._set_observed(observed)


class Discrepancy(NodeReference):
    """discrepancy(*summaries, observed=tuple) -> (B,) or (B, K)."""

    def __init__(self, discrepancy, *parents, **kwargs):
        if not parents:
            raise ValueError('This node requires that at least one parent is specified.')
        state = kwargs.pop('state', None) or {}
        state.update(dict(_operation=discrepancy, _uses_observed=True))
        super().__init__(*parents, state=state, **kwargs)


def _stack_summaries(summaries):
    """np.column_stack(summaries) of elfi/model/utils.py:39 as a device matrix.
    A single 2-d parent is used in place (no copy)."""
    import torch
    cols = []
    for s in summaries:
        t = s if dev.is_device_array(s) else dev.to_device(np.asarray(s, dtype=np.float64))
        if t.dim() > 2:
            raise ValueError('Incompatible data shape for the distance node. Please check '
                             'summary (XA) and observed (XB) output data dimensions. They '
                             'have to be at most 2d.')
        cols.append(t if t.dim() == 2 else t[:, None])
    if len(cols) == 1:
        return cols[0]
    # columns that are adjacent views of one row-major matrix (e.g. the (B, 2) output of the fused
    # MA2 summaries) are re-assembled without a copy
    width = sum(c.shape[1] for c in cols)
    first = cols[0]
    if first.stride(0) == width and all(
            c.stride(0) == width and (c.shape[1] == 1 or c.stride(1) == 1) and
            c.data_ptr() == first.data_ptr() + 8 * sum(x.shape[1] for x in cols[:k]) and
            c.shape[0] == first.shape[0] for k, c in enumerate(cols)):
        return torch.as_strided(first, (first.shape[0], width), (width, 1))
    return torch.cat(cols, dim=1)


_OBSERVED_ROWS = {}     # id(observed tuple) -> [the tuple, stacked host row, its device twin]


def _stack_observed(observed):
    """The observed summaries as one (1, D) host row; one D2H per inference, not per batch."""
    hit = _OBSERVED_ROWS.get(id(observed))
    if hit is not None and hit[0] is observed:
        return hit[1]
    obs = [np.atleast_2d(dev.to_host(o)) for o in observed]
    row = np.concatenate(obs, axis=1).astype(np.float64)
    if len(_OBSERVED_ROWS) > 64:
        _OBSERVED_ROWS.clear()
    _OBSERVED_ROWS[id(observed)] = [observed, row, None]
    return row


def _observed_on_device(observed):
    """Device twin of :func:`_stack_observed` (a pageable H2D copy synchronises the stream: made
    once per inference, the distance kernels of all batches read the same D doubles)."""
    row = _stack_observed(observed)
    hit = _OBSERVED_ROWS[id(observed)]
    if hit[2] is None:
        hit[2] = dev.to_device(row.ravel())
    return hit[2]


_DEVICE_CONSTANTS = {}  # id(host array) -> (the array, device twin): operator constants (w, V)


def _device_constant(arr):
    """Device twin of a host array that an operator was constructed with (cdist's w / V): copied
    once, not with every batch (each pageable H2D copy synchronises the stream)."""
    if arr is None or dev.is_device_array(arr):
        return arr
    hit = _DEVICE_CONSTANTS.get(id(arr))
    if hit is not None and hit[0] is arr:
        return hit[1]
    if len(_DEVICE_CONSTANTS) > 64:
        _DEVICE_CONSTANTS.clear()
    twin = dev.to_device(np.asarray(arr, dtype=np.float64))
    _DEVICE_CONSTANTS[id(arr)] = (arr, twin)
    return twin


def device_euclidean_discrepancy(*summaries, observed, w=None, accept=None):
    """distance_as_discrepancy (elfi/model/utils.py:37-52) for the Euclidean family, on device."""
    X = _stack_summaries(summaries)
    if _stack_observed(observed).shape[0] != 1:
        raise ValueError('observed summaries must form a single row')
    d, idx = ops.dist_euclid(X, _observed_on_device(observed), w=_device_constant(w),
                             thresholds=accept)
    return AcceptedOutput(d, idx) if accept is not None else d


DEVICE_METRICS = ('sqeuclidean', 'cityblock', 'chebyshev', 'minkowski')


def device_metric_discrepancy(metric, *summaries, observed, p=2.0, accept=None):
    """distance_as_discrepancy for the other unweighted cdist metrics that have a kernel."""
    X = _stack_summaries(summaries)
    if _stack_observed(observed).shape[0] != 1:
        raise ValueError('observed summaries must form a single row')
    thr = None if accept is None else np.atleast_1d(dev.to_host(accept))
    d, idx = ops.dist_metric(X, _observed_on_device(observed), metric, p=p, threshold=thr)
    return AcceptedOutput(d, idx) if accept is not None else d


def device_seuclidean_discrepancy(*summaries, observed, V, accept=None):
    """distance_as_discrepancy for cdist's 'seuclidean' (V = component variances)."""
    X = _stack_summaries(summaries)
    if _stack_observed(observed).shape[0] != 1:
        raise ValueError('observed summaries must form a single row')
    thr = None if accept is None else np.atleast_1d(dev.to_host(accept))
    d, idx = ops.dist_seuclidean(X, _observed_on_device(observed), _device_constant(V),
                                 threshold=thr)
    return AcceptedOutput(d, idx) if accept is not None else d


def host_distance_as_discrepancy(dist, *summaries, observed):
    """Generic path for metrics without a CUDA kernel: explicit error, never a silent fallback."""
    raise NotImplementedError(
        "elfi_b200.Distance implements the Euclidean family on the device "
        "('euclidean' with or without w=, 'seuclidean' with V=) and 'sqeuclidean', 'cityblock', "
        "'chebyshev', 'minkowski' (p=) unweighted. Metric {!r} with these keywords has no "
        "CUDA kernel; use elfi_b200.Discrepancy with your own callable.".format(dist))


class Distance(Discrepancy):
    """Distance('euclidean', *summaries[, w=]) -- elfi/model/elfi_model.py:974-1044."""

    def __init__(self, distance, *summaries, **kwargs):
        if not summaries:
            raise ValueError("This node requires that at least one parent is specified.")
        state = {}
        if isinstance(distance, str):
            if distance == 'wminkowski' and 'w' not in kwargs:
                raise ValueError('Parameter w must be specified for distance=wminkowski.')
            if distance == 'seuclidean' and 'V' not in kwargs:
                raise ValueError('Parameter V must be specified for distance=seuclidean.')
            if distance == 'mahalanobis' and 'VI' not in kwargs:
                raise ValueError('Parameter VI must be specified for distance=mahalanobis.')
            cd = {k: kwargs.pop(k) for k in ['p', 'w', 'V', 'VI'] if k in kwargs}
            if distance == 'euclidean' and not (set(cd) - {'w'}):
                op = partial(device_euclidean_discrepancy, w=cd.get('w'))
                state['_uses_accept'] = True
            elif distance == 'seuclidean' and set(cd) == {'V'}:
                op = partial(device_seuclidean_discrepancy,
                             V=np.asarray(cd['V'], dtype=np.float64))
                state['_uses_accept'] = True
            elif distance in DEVICE_METRICS and not (set(cd) - {'p'}) and \
                    (distance == 'minkowski' or not cd):
                op = partial(device_metric_discrepancy, distance, p=cd.get('p', 2.0))
                state['_uses_accept'] = True
            else:
                op = partial(host_distance_as_discrepancy, distance)
        else:
            user_fn = distance

            def op(*summaries, observed):
                X = _stack_summaries(summaries)
                d = user_fn(X, _stack_observed(observed))
                if d.ndim == 2 and d.shape[1] == 1:
                    d = d.reshape(-1)
                return d
        super().__init__(op, *summaries, state=state, **kwargs)
        self.state['attr_dict']['distance'] = distance


class AdaptiveDistance(Discrepancy):
    """Euclidean distance with adaptive per-summary scale (Prangle 2017);
    elfi/model/elfi_model.py:1047-1151.  State: w (list of weight vectors, first None),
    store = [n, mean, M2] merged batch by batch from device column moments."""

    def __init__(self, *summaries, **kwargs):
        if not summaries:
            raise ValueError("This node requires that at least one parent is specified.")
        state = dict(_uses_accept=True)
        super().__init__(self._nested_discrepancy, *summaries, state=state, **kwargs)
        self.init_state()

    # the operation is a bound method of a reference; look the state up at call time
    def _nested_discrepancy(self, *summaries, observed, accept=None):
        X = _stack_summaries(summaries)
        ws = self.state['attr_dict']['w']
        D = X.shape[1]
        key = tuple(id(w) for w in ws)       # (K, D) squared weights: rebuilt when a round is added
        held = self._s.get('_W_dev')
        if held is None or held[0] != key or held[1].shape[1] != D:
            W = np.stack([np.ones(D) if w is None else np.asarray(w, dtype=np.float64) ** 2
                          for w in ws])
            held = self._s['_W_dev'] = (key, dev.to_device(W))
        # the batch's column moments come out of the same read of X; add_data picks them up
        d, idx, mom = ops.dist_euclid(X, _observed_on_device(observed), w=held[1],
                                      thresholds=accept, moments=True)
        self._s['_batch_moments'] = (X, mom)
        return AcceptedOutput(d, idx) if accept is not None else d

    @property
    def _s(self):
        return self.state['attr_dict']

    def init_state(self):
        self._s['w'] = [None]
        self._s['store'] = 3 * [None]
        self.init_adaptation_round()

    def init_adaptation_round(self):
        if 'store' not in self._s:
            self.init_state()
        self._s['store'] = [0, 0, 0]
        self._s.pop('_batch_moments', None)

    def add_data(self, *data):
        """Chan-merge this batch's device column moments into (n, mean, M2); algebraically the
        batch Welford update of elfi_model.py:1117-1123."""
        X = _stack_summaries(data)
        nb = X.shape[0]
        held = self._s.pop('_batch_moments', None)
        if held is not None and held[0].data_ptr() == X.data_ptr() and \
                held[0].shape == X.shape and held[0].stride() == X.stride():
            mean_b, m2_b = held[1].cpu().numpy()      # fused with the distance pass
        else:
            mean_b, m2_b = ops.colmoments(X)
        n0, m0, s0 = self._s['store']
        n1 = n0 + nb
        delta = mean_b - m0
        self._s['store'] = [n1, m0 + delta * (nb / n1), s0 + m2_b + delta ** 2 * (n0 * nb / n1)]
        self._s['scale'] = np.sqrt(self._s['store'][2] / n1)

    def update_distance(self):
        weis = 1 / self._s['scale']
        self._s['w'].append(weis)
        self.init_adaptation_round()

    def nested_distance(self, u, v):
        return self._nested_discrepancy(u, observed=(v,))
