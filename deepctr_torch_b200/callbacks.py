"""Keras-style training callbacks without the tensorflow dependency.

The reference re-exports ``tensorflow.python.keras.callbacks.{EarlyStopping, History,
ModelCheckpoint}`` (reference ``deepctr_torch/callbacks.py:1-73``) and drives them through a
``CallbackList`` from ``BaseModel.fit`` (reference ``deepctr_torch/models/basemodel.py:219-227,
232,303-307``).  This module provides the same classes with the same constructor arguments and
the same observable behaviour (monitor/mode/min_delta/patience semantics, ``History.history``
layout, ``ModelCheckpoint`` file naming and ``torch.save`` payloads) so ``fit(callbacks=[...])``
is a drop-in, with no tensorflow import anywhere.
"""
from __future__ import annotations

import numpy as np
import torch


class Callback:
    """Base class: every hook is a no-op; ``model`` / ``params`` are injected by CallbackList."""

    def __init__(self):
        self.model = None
        self.params = None
        self.validation_data = None

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


class CallbackList:
    """Fan-out container used by ``fit`` (reference basemodel.py:219-227)."""

    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])
        self.model = None
        self.params = None

    def append(self, callback):
        self.callbacks.append(callback)

    def set_params(self, params):
        self.params = params
        for cb in self.callbacks:
            cb.set_params(params)

    def set_model(self, model):
        self.model = model
        for cb in self.callbacks:
            cb.set_model(model)

    def on_train_begin(self, logs=None):
        for cb in self.callbacks:
            cb.on_train_begin(logs)

    def on_train_end(self, logs=None):
        for cb in self.callbacks:
            cb.on_train_end(logs)

    def on_epoch_begin(self, epoch, logs=None):
        for cb in self.callbacks:
            cb.on_epoch_begin(epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        for cb in self.callbacks:
            cb.on_epoch_end(epoch, logs)

    def on_batch_begin(self, batch, logs=None):
        for cb in self.callbacks:
            cb.on_batch_begin(batch, logs)

    def on_batch_end(self, batch, logs=None):
        for cb in self.callbacks:
            cb.on_batch_end(batch, logs)

    def __iter__(self):
        return iter(self.callbacks)


class History(Callback):
    """Records per-epoch logs into ``history`` (dict name -> list) and ``epoch`` (list)."""

    def __init__(self):
        super().__init__()
        self.history = {}
        self.epoch = []

    def on_train_begin(self, logs=None):
        self.epoch = []

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epoch.append(epoch)
        for key, value in logs.items():
            self.history.setdefault(key, []).append(value)


def _resolve_monitor_op(mode, monitor):
    if mode not in ("auto", "min", "max"):
        mode = "auto"
    if mode == "min":
        return np.less
    if mode == "max":
        return np.greater
    if "acc" in monitor or monitor.startswith("fmeasure") or "auc" in monitor:
        return np.greater
    return np.less


class EarlyStopping(Callback):
    """Stop training when ``monitor`` has stopped improving (Keras semantics)."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto",
                 baseline=None, restore_best_weights=False):
        super().__init__()
        self.monitor = monitor
        self.patience = patience
        self.verbose = verbose
        self.baseline = baseline
        self.min_delta = abs(min_delta)
        self.wait = 0
        self.stopped_epoch = 0
        self.restore_best_weights = restore_best_weights
        self.best_weights = None
        self.monitor_op = _resolve_monitor_op(mode, monitor)
        if self.monitor_op == np.greater:
            self.min_delta *= 1
        else:
            self.min_delta *= -1
        self.best = None

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        if self.baseline is not None:
            self.best = self.baseline
        else:
            self.best = np.inf if self.monitor_op == np.less else -np.inf

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            return
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
            if self.restore_best_weights:
                self.best_weights = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore_best_weights and self.best_weights is not None:
                    self.model.load_state_dict(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose > 0:
            print("Epoch %05d: early stopping" % (self.stopped_epoch + 1))


class ModelCheckpoint(Callback):
    """Save the model (``torch.save``) after every ``period`` epochs.

    Same arguments as the reference's subclass of the Keras callback
    (reference ``deepctr_torch/callbacks.py:9-73``): ``save_weights_only`` stores
    ``model.state_dict()``, otherwise the whole module is pickled.
    """

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False,
                 save_weights_only=False, mode="auto", period=1):
        super().__init__()
        self.monitor = monitor
        self.verbose = verbose
        self.filepath = filepath
        self.save_best_only = save_best_only
        self.save_weights_only = save_weights_only
        self.period = period
        self.epochs_since_last_save = 0
        self.monitor_op = _resolve_monitor_op(mode, monitor)
        self.best = np.inf if self.monitor_op == np.less else -np.inf

    def _save(self, filepath):
        if self.save_weights_only:
            torch.save(self.model.state_dict(), filepath)
        else:
            torch.save(self.model, filepath)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        filepath = self.filepath.format(epoch=epoch + 1, **logs)
        if not self.save_best_only:
            if self.verbose > 0:
                print("Epoch %05d: saving model to %s" % (epoch + 1, filepath))
            self._save(filepath)
            return
        current = logs.get(self.monitor)
        if current is None:
            print("Can save best model only with %s available, skipping." % self.monitor)
        elif self.monitor_op(current, self.best):
            if self.verbose > 0:
                print("Epoch %05d: %s improved from %0.5f to %0.5f, saving model to %s"
                      % (epoch + 1, self.monitor, self.best, current, filepath))
            self.best = current
            self._save(filepath)
        elif self.verbose > 0:
            print("Epoch %05d: %s did not improve from %0.5f" % (epoch + 1, self.monitor, self.best))
