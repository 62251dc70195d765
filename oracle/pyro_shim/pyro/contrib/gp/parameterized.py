from collections import OrderedDict

from pyro.nn import PyroModule


class Parameterized(PyroModule):
    """mode switch + prior/guide bookkeeping; the reference sets no priors, so ``_load_pyro_samples`` finds nothing."""

    def __init__(self):
        super().__init__()
        self._priors = OrderedDict()
        self._guides = OrderedDict()
        self._mode = "model"

    def set_prior(self, name, prior):
        raise NotImplementedError("shim: priors on parameters are not on the accelerated path")

    def autoguide(self, name, dist_constructor):
        raise NotImplementedError("shim: autoguides are not on the accelerated path")

    def set_mode(self, mode):
        for module in self.modules():
            if isinstance(module, Parameterized):
                module.mode = mode

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, mode):
        self._mode = mode

    def _load_pyro_samples(self):
        for module in self.modules():
            if isinstance(module, Parameterized):
                for name in module._priors:
                    getattr(module, name)
