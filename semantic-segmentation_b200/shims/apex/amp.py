import contextlib


def float_function(fn):
    return fn


def half_function(fn):
    return fn


@contextlib.contextmanager
def disable_casts():
    yield


def initialize(models, optimizers=None, opt_level="O1", **kwargs):
    if optimizers is None:
        return models
    return models, optimizers


@contextlib.contextmanager
def scale_loss(loss, optimizers, **kwargs):
    yield loss
