"""Eval-mode forward (placeholder until the device-side blend/argmax kernels land)."""


def eval_forward(module, images):
    raise NotImplementedError("eval-mode forward of the B200 path is not implemented yet")
