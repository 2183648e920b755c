"""Optimizer spec objects with the constructor surface of `keras.optimizers` as the reference uses it
(experiments_methods.py:10, :41: `Adagrad(lr=lr, epsilon=1e-08, decay=0.0, clipnorm=1.)`).  They only carry
hyper-parameters; the arithmetic is the K8 kernels (csrc/optim.cu)."""


class Optimizer(object):
    kind = None

    def __init__(self, clipnorm=None, clipvalue=None):
        if clipvalue is not None:
            raise NotImplementedError("clipvalue is never used by the reference")
        self.clipnorm = clipnorm


class Adagrad(Optimizer):
    kind = "adagrad"

    def __init__(self, lr=0.01, epsilon=1e-8, decay=0.0, **kwargs):
        Optimizer.__init__(self, **kwargs)
        self.lr = lr
        self.epsilon = epsilon
        self.decay = decay


def resolve(optimizer):
    """Accept an optimizer object (anything exposing lr / epsilon / decay / clipnorm, e.g. a Keras Adagrad) or the
    name 'adagrad'.  `compile_model`'s default string 'adam' (model.py:176) is accepted at compile time but the
    reference never trains with it; fitting with it raises."""
    if isinstance(optimizer, str):
        if optimizer.lower() == "adagrad":
            return Adagrad()
        return optimizer.lower()
    kind = getattr(optimizer, "kind", type(optimizer).__name__.lower())
    if kind != "adagrad":
        return kind
    return optimizer
