"""trace / replay / scale, as far as Trace_ELBO over the reference's model and guide needs them."""
from collections import OrderedDict

import pyro


class Trace:
    def __init__(self):
        self.nodes = OrderedDict()


class _TraceHandler(pyro._Handler):
    def __init__(self):
        self.trace = Trace()

    def postprocess(self, msg):
        if msg["name"] in self.trace.nodes:
            raise RuntimeError(f"Multiple sample sites named '{msg['name']}'")
        self.trace.nodes[msg["name"]] = dict(msg)


class _ReplayHandler(pyro._Handler):
    def __init__(self, trace):
        self.guide_trace = trace

    def process(self, msg):
        if not msg["is_observed"] and msg["name"] in self.guide_trace.nodes:
            msg["value"] = self.guide_trace.nodes[msg["name"]]["value"]


class _ScaleHandler(pyro._Handler):
    def __init__(self, scale):
        self.scale = scale

    def process(self, msg):
        msg["scale"] = msg["scale"] * self.scale

    def __call__(self, fn):      # handler used as a decorator: scale = poutine.scale(scale=1/N); scale(model)  (train_script.py:365)
        return _Wrapped(fn, lambda: _ScaleHandler(self.scale))


class _Wrapped:
    def __init__(self, fn, make_handler):
        self.fn, self.make_handler = fn, make_handler

    def __call__(self, *args, **kwargs):
        with self.make_handler():
            return self.fn(*args, **kwargs)


class _Tracer(_Wrapped):
    def __init__(self, fn):
        super().__init__(fn, None)

    def get_trace(self, *args, **kwargs):
        h = _TraceHandler()
        with h:
            self.fn(*args, **kwargs)
        return h.trace


def trace(fn):
    return _Tracer(fn)


def replay(fn, trace=None):
    return _Wrapped(fn, lambda: _ReplayHandler(trace))


def scale(fn=None, scale=1.0):
    if fn is None:
        return _ScaleHandler(scale)
    return _Wrapped(fn, lambda: _ScaleHandler(scale))
