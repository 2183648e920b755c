"""Callback protocol of Keras-2.0.x as far as the reference uses it (experiments_methods.py:22-38, model.py:94-117):
`Callback` with a `.model` back-reference, `History`, `EarlyStopping(monitor, min_delta, patience)`,
`ModelCheckpoint(filepath, monitor, save_weights_only, save_best_only)`."""
import warnings

import numpy as np


class Callback(object):
    def __init__(self):
        self.validation_data = None
        self.model = None
        self.params = {}

    def set_params(self, params):
        self.params = params

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass


class History(Callback):
    def on_train_begin(self, logs=None):
        self.epoch = []
        self.history = {}

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


def _monitor_op(mode, monitor):
    if mode == "max" or (mode == "auto" and "acc" in monitor):
        return np.greater, -np.inf
    return np.less, np.inf


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto"):
        Callback.__init__(self)
        self.monitor, self.patience, self.verbose = monitor, patience, verbose
        self.monitor_op, self._init = _monitor_op(mode, monitor)
        self.min_delta = min_delta if self.monitor_op == np.greater else -min_delta
        self.wait = 0
        self.stopped_epoch = 0

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        self.best = self._init

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            warnings.warn("Early stopping requires %s available!" % self.monitor, RuntimeWarning)
            return
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
        else:
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
            self.wait += 1


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        Callback.__init__(self)
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.save_weights_only, self.period = save_best_only, save_weights_only, period
        self.epochs_since_last_save = 0
        self.monitor_op, self.best = _monitor_op(mode, monitor)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        filepath = self.filepath.format(epoch=epoch, **logs)
        if self.save_best_only:
            current = logs.get(self.monitor)
            if current is None:
                warnings.warn("Can save best model only with %s available, skipping." % self.monitor, RuntimeWarning)
                return
            if not self.monitor_op(current, self.best):
                return
            self.best = current
        self.model.save_weights(filepath, overwrite=True)
