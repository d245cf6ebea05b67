"""Small helpers on the sampling path (reference: k_diffusion/utils.py:43-48,82-85)."""
from contextlib import contextmanager


def append_dims(x, target_dims):
    """Right-pad x's shape with singleton dims up to `target_dims` (utils.py:43-48)."""
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f'input has {x.ndim} dims but target_dims is {target_dims}, which is less')
    return x.reshape(tuple(x.shape) + (1,) * extra)


def n_params(module):
    return sum(p.numel() for p in module.parameters())


@contextmanager
def _mode(model, training):
    was = [m.training for m in model.modules()]
    try:
        yield model.train(training)
    finally:
        for m, t in zip(model.modules(), was):
            m.training = t


def eval_mode(model):
    """Context manager: put `model` in eval mode, restore on exit (utils.py:82)."""
    return _mode(model, False)


def train_mode(model):
    return _mode(model, True)
