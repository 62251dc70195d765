"""pyro.optim.X({"lr": ...}): one torch optimiser instance per parameter, created on first sight."""
import torch


class PyroOptim:
    def __init__(self, optim_constructor, optim_args):
        self.ctor, self.args, self.optim_objs = optim_constructor, optim_args, {}

    def __call__(self, params):
        for p in params:
            if p not in self.optim_objs:
                self.optim_objs[p] = self.ctor([p], **self.args)
            self.optim_objs[p].step()


def Adam(optim_args):
    return PyroOptim(torch.optim.Adam, optim_args)


def AdamW(optim_args):
    return PyroOptim(torch.optim.AdamW, optim_args)


def SGD(optim_args):
    return PyroOptim(torch.optim.SGD, optim_args)
