"""Host-side logic of the reference-facing surface that needs no GPU: the Keras-2.0.x callback protocol the reference's
training drivers rely on (experiments_methods.py:30-38: EarlyStopping(monitor, min_delta=0, patience=15, mode='auto'),
ModelCheckpoint(filepath with {epoch:02d}-{val_loss:.2f}, save_best_only=True)), the History object `fit_model` returns
(model.py:179-182; experiments_methods.py:92-93 reads .history['loss'] / ['val_loss']) and the optimizer spec objects
(experiments_methods.py:41)."""
import warnings

import numpy as np
import pytest

from seq_recommendations_b200 import callbacks as cb
from seq_recommendations_b200 import optimizers as opt


class FakeNet(object):
    def __init__(self):
        self.stop_training = False
        self.saved = []

    def save_weights(self, filepath, overwrite=True):
        self.saved.append(filepath)


def run(callback, values, monitor="val_loss"):
    """Drive one callback through a training run the way `_Net.fit` does; returns the epochs that ran."""
    net = FakeNet()
    callback.set_model(net)
    callback.on_train_begin({})
    ran = []
    for epoch, v in enumerate(values):
        callback.on_epoch_begin(epoch, {})
        callback.on_epoch_end(epoch, {monitor: v, "loss": v})
        ran.append(epoch)
        if net.stop_training:
            break
    callback.on_train_end({})
    return net, ran


def test_early_stopping_patience_counts_epochs_without_improvement():
    # Keras 2.0.x: an epoch that does not improve first checks `wait >= patience`, then increments -- so with patience p
    # training stops at the (p+1)-th consecutive epoch without improvement
    stop = cb.EarlyStopping(monitor="val_loss", min_delta=0, patience=2, mode="auto")
    net, ran = run(stop, [1.0, 0.9, 0.95, 0.91, 0.92, 0.5, 0.4])
    assert ran == [0, 1, 2, 3, 4] and stop.stopped_epoch == 4 and stop.best == 0.9
    # an improvement resets the count
    stop = cb.EarlyStopping(monitor="val_loss", patience=2)
    net, ran = run(stop, [1.0, 1.1, 1.2, 0.8, 0.9, 1.0, 1.1])
    assert ran == [0, 1, 2, 3, 4, 5, 6] and stop.stopped_epoch == 6 and stop.best == 0.8
    # patience 0 stops at the first epoch that is not better; equal is not better
    stop = cb.EarlyStopping(monitor="val_loss", patience=0)
    net, ran = run(stop, [1.0, 1.0, 0.5])
    assert ran == [0, 1]
    # a second fit starts from scratch
    net, ran = run(stop, [3.0, 2.0, 1.0])
    assert ran == [0, 1, 2] and stop.best == 1.0


def test_early_stopping_min_delta_and_modes():
    stop = cb.EarlyStopping(monitor="val_loss", min_delta=0.1, patience=0)
    net, ran = run(stop, [1.0, 0.95, 0.5])               # 0.95 is not better than 1.0 by more than 0.1
    assert ran == [0, 1]
    stop = cb.EarlyStopping(monitor="val_acc", patience=0, mode="auto")      # 'acc' in the name -> larger is better
    net, ran = run(stop, [0.1, 0.2, 0.15, 0.9], monitor="val_acc")
    assert ran == [0, 1, 2] and stop.best == 0.2
    stop = cb.EarlyStopping(monitor="my_loss", patience=0, mode="max")
    net, ran = run(stop, [0.1, 0.2, 0.15], monitor="my_loss")
    assert ran == [0, 1, 2]
    stop = cb.EarlyStopping(monitor="my_loss", patience=1)                   # the reference's second monitor name
    net, ran = run(stop, [2.0, 1.0, 1.5, 1.6, 0.1], monitor="my_loss")
    assert ran == [0, 1, 2, 3]


def test_early_stopping_warns_when_the_monitored_quantity_is_missing():
    stop = cb.EarlyStopping(monitor="val_loss", patience=0)
    stop.set_model(FakeNet())
    stop.on_train_begin({})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        stop.on_epoch_end(0, {"loss": 1.0})
    assert len(w) == 1 and issubclass(w[0].category, RuntimeWarning)
    assert not stop.model.stop_training


def test_model_checkpoint_save_best_only_and_filename_template(tmp_path):
    tmpl = str(tmp_path / "weights.{epoch:02d}-{val_loss:.2f}.npz")
    ckpt = cb.ModelCheckpoint(tmpl, monitor="val_loss", save_best_only=True, save_weights_only=True, mode="auto")
    net, ran = run(ckpt, [1.0, 1.25, 0.75, 0.75, 0.5])
    assert net.saved == [str(tmp_path / "weights.00-1.00.npz"), str(tmp_path / "weights.02-0.75.npz"),
                         str(tmp_path / "weights.04-0.50.npz")]
    ckpt = cb.ModelCheckpoint(tmpl, monitor="val_loss", save_best_only=False, period=2)
    net, ran = run(ckpt, [1.0, 2.0, 3.0, 4.0, 5.0])
    assert net.saved == [str(tmp_path / "weights.01-2.00.npz"), str(tmp_path / "weights.03-4.00.npz")]
    ckpt = cb.ModelCheckpoint(tmpl, monitor="val_acc", save_best_only=True)
    ckpt.set_model(FakeNet())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ckpt.on_epoch_end(0, {"val_loss": 1.0})
    assert ckpt.model.saved == [] and len(w) == 1


def test_history_collects_per_epoch_lists():
    h = cb.History()
    h.set_model(FakeNet())
    h.on_train_begin({})
    for e, (l, v) in enumerate([(3.0, 3.5), (2.0, 2.5), (1.5, 2.75)]):
        h.on_epoch_end(e, {"loss": l, "val_loss": v})
    assert h.epoch == [0, 1, 2]
    assert h.history == {"loss": [3.0, 2.0, 1.5], "val_loss": [3.5, 2.5, 2.75]}
    # what experiments_methods.analyze_history does with it (:91-97): position and value of the best validation loss
    best = int(np.argmin(h.history["val_loss"]))
    assert (best, h.history["val_loss"][best], h.history["loss"][best]) == (1, 2.5, 2.0)
    h.on_train_begin({})
    assert h.epoch == [] and h.history == {}


def test_optimizer_spec_objects():
    o = opt.Adagrad(lr=0.05, epsilon=1e-08, decay=0.0, clipnorm=1.)          # experiments_methods.py:41 verbatim
    assert (o.kind, o.lr, o.epsilon, o.decay, o.clipnorm) == ("adagrad", 0.05, 1e-8, 0.0, 1.0)
    assert opt.resolve(o) is o
    d = opt.resolve("adagrad")
    assert isinstance(d, opt.Adagrad) and (d.lr, d.epsilon, d.decay, d.clipnorm) == (0.01, 1e-8, 0.0, None)
    assert opt.resolve("adam") == "adam"                                     # compile_model's default string: accepted
    with pytest.raises(NotImplementedError):
        opt.Adagrad(clipvalue=0.5)

    class SGD(object):                                                       # a foreign optimizer object: named, not run
        lr = 0.1
    assert opt.resolve(SGD()) == "sgd"

    class KerasLikeAdagrad(object):                                          # e.g. a real keras.optimizers.Adagrad
        kind = "adagrad"
        lr, epsilon, decay, clipnorm = 0.1, 1e-8, 0.0, 1.0
    k = KerasLikeAdagrad()
    assert opt.resolve(k) is k
