"""TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the TensorFlow-1.x API surface the reference's model files use,
backed by torch CPU ops + autograd, so that the reference's OWN graph-construction code (src/models/pl/models/bprmf.py,
cml.py, gbprmf.py, prigp.py, cplr_u.py, src/models/basic/models/wrmf.py, mf.py, svd.py) can be imported unmodified from
/root/reference and executed in this container (TensorFlow itself has no wheel here and there is no network).  Only
oracle/gen_refgraph_golden.py imports it (with this directory put first on sys.path); nothing in the product does.

What is emulated, and the TF-1.x semantics it follows:
  * a deferred graph: every tf.* call returns a Node; Session.run evaluates the fetched structure with a per-run memo.
    Nodes created inside ``tf.control_dependencies(ops)`` evaluate ``ops`` first.  Fetches that do not depend on a
    stateful op (minimize / assign) are evaluated BEFORE the stateful ones: the loss fetched next to the train op is
    the pre-update loss (in TF1 the order is formally undefined; the gather ops of the separately built loss graph are
    ready at once while the apply ops wait for a whole forward + backward, SURVEY D-list).
  * Variables are ref variables: a consumer sees the value at the time IT runs (so CML's clip ops, created under the
    control dependency on the train op, clip the updated tables: cml.py:119-129).
  * tf.train.AdagradOptimizer(lr).minimize: initial_accumulator_value 0.1, one accumulator set per optimizer INSTANCE
    (each evaluation of the reference's ``__optimize__`` property builds a new one), acc += g^2, var -= lr * g / sqrt(acc).
    Gradients of embedding_lookup are IndexedSlices in TF; optimizer.py::_apply_sparse_duplicate_indices sums duplicate
    indices before the apply and leaves untouched rows alone -- which is what a dense autograd gradient does too
    (g = 0: acc += 0, var -= 0), so the dense form is used.
  * reduce_min's gradient is shared equally between tied minima (math_grad.py::_MinOrMaxGrad) = torch.amin's.
  * relu'(0) = 0; clip_by_norm = t * clip / max(||t||, clip) along ``axes``; l2_loss = sum(t^2) / 2;
    top_k breaks ties towards the lower index.
  * initializers draw from numpy (seeded by ``set_random_seed``) unless ``INIT_OVERRIDE[name]`` holds the array to use.
Arithmetic is float32 on one thread.  This is an interpretation of TensorFlow, not TensorFlow: DESIGN.md section 2 says so."""
import contextlib

import numpy as np
import torch

float32, int32, int64 = torch.float32, torch.int32, torch.int64
torch.set_num_threads(1)

INIT_OVERRIDE = {}       # variable name -> numpy array used by global_variables_initializer instead of the initializer
VARIABLES = []           # every Variable created since reset_default_graph()
OPTIMIZERS = []          # every AdagradOptimizer created since reset_default_graph()
_CONTROL = []            # stack of control-dependency lists
_RNG = [np.random.default_rng(0)]


def reset_default_graph():
    del VARIABLES[:], OPTIMIZERS[:], _CONTROL[:]
    INIT_OVERRIDE.clear()


def set_random_seed(seed):
    _RNG[0] = np.random.default_rng(seed)


class Node(object):
    stateful = False

    def __init__(self, fn, inputs=(), name=None):
        self.fn, self.inputs, self.name = fn, [_wrap(x) for x in inputs], name
        self.control = [c for deps in _CONTROL for c in deps]

    def _eval(self, run):
        if id(self) in run.memo:
            return run.memo[id(self)]
        for c in self.control:
            run.fetch(c)
        v = self.fn(*[x._eval(run) for x in self.inputs])
        run.memo[id(self)] = v
        return v

    # operators the reference uses on tensors
    def __add__(self, o): return Node(lambda a, b: a + b, (self, o))
    __radd__ = __add__
    def __sub__(self, o): return Node(lambda a, b: a - b, (self, o))
    def __rsub__(self, o): return Node(lambda a, b: b - a, (self, o))
    def __mul__(self, o): return Node(lambda a, b: a * b, (self, o))
    __rmul__ = __mul__
    def __truediv__(self, o): return Node(lambda a, b: a / b, (self, o))
    def __neg__(self): return Node(lambda a: -a, (self,))
    def __gt__(self, o): return Node(lambda a, b: a > b, (self, o))
    def __getitem__(self, idx): return Node(lambda a: a[idx], (self,))
    __hash__ = object.__hash__


class Const(Node):
    def __init__(self, value):
        Node.__init__(self, None)
        self.value = value

    def _eval(self, run):
        return self.value


def _wrap(x):
    if isinstance(x, Node):
        return x
    if isinstance(x, (np.floating, float, int, np.integer)):
        return Const(torch.tensor(float(x), dtype=torch.float32) if isinstance(x, (float, np.floating)) else int(x))
    return Const(torch.as_tensor(np.asarray(x)))


class Placeholder(Node):
    def __init__(self, dtype, shape=None, name=None):
        Node.__init__(self, None, name=name)
        self.dtype = dtype

    def _eval(self, run):
        if self not in run.feed:
            raise KeyError('placeholder %r was not fed' % self.name)
        v = torch.as_tensor(np.asarray(run.feed[self]))
        return v.to(torch.float32) if self.dtype == float32 else v.to(torch.int64)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


class Variable(Node):
    def __init__(self, name, shape, initializer):
        Node.__init__(self, None, name=name)
        self.shape, self.initializer, self.value = [int(s) for s in shape], initializer, None
        VARIABLES.append(self)

    def _eval(self, run):
        if self.value is None:
            raise RuntimeError('variable %s is not initialised' % self.name)
        return self.value            # the live tensor: a consumer sees the value at the time it runs (ref variable)


def get_variable(name=None, shape=None, initializer=None, dtype=None):
    return Variable(name, shape, initializer)


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    def draw(shape):
        x = _RNG[0].standard_normal(shape)
        bad = np.abs(x) > 2.0
        while bad.any():                                  # values beyond two standard deviations are re-drawn
            x[bad] = _RNG[0].standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (mean + stddev * x).astype(np.float32)
    return draw


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    return lambda shape: (mean + stddev * _RNG[0].standard_normal(shape)).astype(np.float32)


class _Stateful(Node):
    stateful = True


def global_variables_initializer():
    def init():
        for v in VARIABLES:
            a = INIT_OVERRIDE[v.name] if v.name in INIT_OVERRIDE else v.initializer(v.shape)
            a = np.asarray(a, dtype=np.float32)
            assert list(a.shape) == v.shape, (v.name, a.shape, v.shape)
            v.value = torch.tensor(a.copy(), requires_grad=True)
    return _Stateful(init)


def assign(ref, value, name=None):
    def f(val):
        with torch.no_grad():
            ref.value.copy_(val)
        return ref.value
    return _Stateful(f, (value,), name)


@contextlib.contextmanager
def control_dependencies(ops):
    _CONTROL.append(list(ops))
    try:
        yield
    finally:
        _CONTROL.pop()


@contextlib.contextmanager
def device(name):
    yield


# ---- math
def _axis(axis, reduction_indices):
    a = reduction_indices if reduction_indices is not None else axis
    return tuple(a) if isinstance(a, (list, tuple)) else a


def reduce_sum(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    a = _axis(axis, reduction_indices)
    return Node((lambda t: t.sum()) if a is None else (lambda t: t.sum(dim=a, keepdim=keepdims)), (x,), name)


def reduce_mean(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    a = _axis(axis, reduction_indices)
    return Node((lambda t: t.mean()) if a is None else (lambda t: t.mean(dim=a, keepdim=keepdims)), (x,), name)


def reduce_min(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    a = _axis(axis, reduction_indices)
    return Node((lambda t: t.amin()) if a is None else (lambda t: t.amin(dim=a, keepdim=keepdims)), (x,), name)


def expand_dims(x, axis=None, name=None, dim=None):
    a = axis if axis is not None else dim
    return Node(lambda t: t.unsqueeze(a), (x,), name)


def add(x, y, name=None): return Node(lambda a, b: a + b, (x, y), name)
def divide(x, y, name=None): return Node(lambda a, b: a / b, (x, y), name)
def multiply(x, y, name=None): return Node(lambda a, b: a * b, (x, y), name)
def subtract(x, y, name=None): return Node(lambda a, b: a - b, (x, y), name)
def squared_difference(x, y, name=None): return Node(lambda a, b: (a - b) * (a - b), (x, y), name)
def log(x, name=None): return Node(torch.log, (x,), name)
def sigmoid(x, name=None): return Node(torch.sigmoid, (x,), name)
def square(x, name=None): return Node(lambda a: a * a, (x,), name)
def sqrt(x, name=None): return Node(torch.sqrt, (x,), name)
def transpose(x, perm=None, name=None): return Node((lambda a: a.t()) if perm is None else (lambda a: a.permute(*perm)), (x,), name)
def greater(x, y, name=None): return Node(lambda a, b: a > b, (x, y), name)
def less_equal(x, y, name=None): return Node(lambda a, b: a <= b, (x, y), name)
def shape(x, name=None): return Node(lambda a: torch.tensor(list(a.shape)), (x,), name)
def zeros(shape, dtype=float32, name=None): return Const(torch.zeros(shape, dtype=dtype))
def concat(values, axis=0, name=None): return Node(lambda *a: torch.cat(a, dim=axis), tuple(values), name)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    return Node(lambda x, y: (x.t() if transpose_a else x) @ (y.t() if transpose_b else y), (a, b), name)


def cast(x, dtype, name=None):
    if not isinstance(x, Node):
        return Const(torch.tensor(x, dtype=dtype))
    return Node(lambda a: a.to(dtype), (x,), name)


def clip_by_norm(t, clip_norm, axes=None, name=None):
    def f(a):                                             # clip_ops.py: t * clip / max(l2norm, clip)
        l2sum = (a * a).sum(dim=tuple(axes) if axes is not None else None, keepdim=True)
        pred = l2sum > 0
        l2norm = torch.where(pred, torch.sqrt(torch.where(pred, l2sum, torch.ones_like(l2sum))), l2sum)
        return (a * clip_norm) / torch.maximum(l2norm, torch.tensor(float(clip_norm)))
    return Node(f, (t,), name)


class _TopK(tuple):
    values = property(lambda s: s[0])
    indices = property(lambda s: s[1])


class nn(object):
    @staticmethod
    def embedding_lookup(params, ids, name=None):
        return Node(lambda p, i: p[i.long()], (params, ids), name)

    @staticmethod
    def l2_loss(t, name=None):
        return Node(lambda a: (a * a).sum() / 2, (t,), name)

    @staticmethod
    def relu(x, name=None):
        return Node(torch.relu, (x,), name)

    @staticmethod
    def top_k(x, k=1, sorted=True, name=None):
        def f(a):                                         # ties towards the lower index, like TopKV2
            order = torch.argsort(-a.detach(), dim=-1, stable=True)[..., :int(k)]
            return torch.gather(a.detach(), -1, order), order.to(torch.int32)
        both = Node(f, (x,), name)
        return _TopK((Node(lambda b: b[0], (both,)), Node(lambda b: b[1], (both,))))


class _Adagrad(object):
    def __init__(self, learning_rate, initial_accumulator_value=0.1, name='Adagrad'):
        self.lr, self.init_acc, self.accum = float(learning_rate), float(initial_accumulator_value), {}
        OPTIMIZERS.append(self)

    def minimize(self, loss, var_list=None, name=None):
        opt = self
        var_list = list(var_list) if var_list is not None else list(VARIABLES)
        for v in var_list:                                 # slots are created when the op is built
            opt.accum[v] = None

        class Minimize(_Stateful):
            def _eval(self, run):
                if id(self) in run.memo:
                    return None
                for c in self.control:
                    run.fetch(c)
                with torch.enable_grad():
                    sub = _Run(run.feed)                   # the loss of THIS op's own graph, under autograd
                    value = loss._eval(sub)
                    grads = torch.autograd.grad(value, [v.value for v in var_list], allow_unused=True)
                with torch.no_grad():
                    for v, g in zip(var_list, grads):
                        if opt.accum[v] is None:
                            opt.accum[v] = torch.full_like(v.value, opt.init_acc)
                        if g is None:
                            continue
                        opt.accum[v] += g * g
                        v.value -= opt.lr * g / torch.sqrt(opt.accum[v])
                run.memo[id(self)] = True
                return None
        return Minimize(None, (), name)


class train(object):
    AdagradOptimizer = _Adagrad


# ---- session
def _depends_on_state(node, seen):
    if id(node) in seen:
        return seen[id(node)]
    seen[id(node)] = False
    r = node.stateful or any(_depends_on_state(x, seen) for x in list(node.inputs) + list(node.control))
    seen[id(node)] = r
    return r


class _Run(object):
    def __init__(self, feed):
        self.feed, self.memo = feed, {}

    def fetch(self, f):
        if isinstance(f, (list, tuple)):
            return [self.fetch(x) for x in f]
        v = f._eval(self)
        if torch.is_tensor(v):
            return v.detach().numpy().copy()
        return v


def _leaves(f, out):
    if isinstance(f, (list, tuple)):
        for x in f:
            _leaves(x, out)
    else:
        out.append(f)
    return out


class _GpuOptions(object):
    allow_growth = False


class ConfigProto(object):
    def __init__(self, **kw):
        self.gpu_options = _GpuOptions()


class Session(object):
    def __init__(self, config=None, **kw):
        pass

    def run(self, fetches, feed_dict=None):
        run = _Run(dict(feed_dict or {}))
        with torch.no_grad():
            seen = {}
            pure = {}
            for leaf in _leaves(fetches, []):              # phase 1: fetches that no stateful op feeds (e.g. the loss)
                if not _depends_on_state(leaf, seen):
                    pure[id(leaf)] = run.fetch(leaf)

            def second(f):
                if isinstance(f, (list, tuple)):
                    return [second(x) for x in f]
                return pure[id(f)] if id(f) in pure else run.fetch(f)
            out = second(fetches)
        return out if isinstance(fetches, (list, tuple)) else out

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return Node(lambda a: torch.clamp(a, float(clip_value_min), float(clip_value_max)), (t,), name)
