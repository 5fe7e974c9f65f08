"""Minimal `pytorch_lightning`: the LightningModule surface the reference's trainer/trainer.py touches and a plain
single-device fit loop (prepare_data -> configure_optimizers -> train_dataloader -> training_step / backward / step ->
training_epoch_end -> callbacks).  No validation loop, no checkpoint formats beyond a state dict, no distributed strategy: the
data-parallel half of the path is unscene3d_b200/distributed.py."""
import random

import numpy as np
import torch
from torch import nn

from . import callbacks, loggers  # noqa: F401
from .callbacks import Callback

__version__ = "1.7.2+us3d-standin"


def seed_everything(seed=None, workers=False):
    seed = 0 if seed is None else int(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def _move(obj, device):
    """Tensors inside lists / tuples / dicts move; anything else (the reference's NoGpu wrapper) stays as it is."""
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: _move(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_move(v, device) for v in obj)
    return obj


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.trainer = None
        self.current_epoch = 0
        self.global_step = 0
        self.logged = {}
        self.hparams = {}

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def save_hyperparameters(self, *args, **kwargs):
        return None

    def log(self, name, value, *args, **kwargs):
        self.logged[name] = value

    def log_dict(self, dictionary, *args, **kwargs):
        self.logged.update(dictionary)

    # hooks the loop calls when they exist
    def prepare_data(self):
        pass

    def training_epoch_end(self, outputs):
        pass


class Trainer:
    def __init__(self, logger=None, gpus=None, callbacks=None, max_epochs=1, min_epochs=1, max_steps=-1, limit_train_batches=None,
                 weights_save_path=None, resume_from_checkpoint=None, **unused):
        self.logger = logger
        self.callbacks = list(callbacks or [])
        self.max_epochs, self.max_steps, self.limit_train_batches = int(max_epochs), int(max_steps or -1), limit_train_batches
        self.gpus = gpus
        self.weights_save_path = weights_save_path
        self.global_step = 0
        self.current_epoch = 0
        self.model = None
        self.losses = []

    def save_checkpoint(self, path):
        import os

        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"state_dict": self.model.state_dict(), "epoch": self.current_epoch, "global_step": self.global_step}, path)

    def fit(self, model):
        self.model = model
        model.trainer = self
        device = torch.device("cuda", torch.cuda.current_device()) if (self.gpus and torch.cuda.is_available()) else torch.device("cpu")
        model.to(device)
        model.prepare_data()
        optimizers, schedulers = model.configure_optimizers()
        optimizer = optimizers[0]
        sched = schedulers[0] if schedulers else None
        loader = model.train_dataloader()
        limit = self.limit_train_batches
        done = False
        for epoch in range(self.max_epochs):
            self.current_epoch = model.current_epoch = epoch
            model.train()
            outputs = []
            for i, batch in enumerate(loader):
                if limit is not None and i >= (int(limit) if limit >= 1 else max(int(limit * len(loader)), 1)):
                    break
                batch = _move(batch, device)
                loss = model.training_step(batch, i)
                if loss is None:
                    continue
                optimizer.zero_grad(set_to_none=True)
                loss.backward()
                optimizer.step()
                if sched is not None and sched.get("interval", "epoch") == "step":
                    sched["scheduler"].step()
                self.global_step += 1
                model.global_step = self.global_step
                self.losses.append(float(loss.detach()))
                outputs.append({"loss": loss.detach()})
                if 0 < self.max_steps <= self.global_step:
                    done = True
                    break
            if outputs:
                model.training_epoch_end(outputs)
            if sched is not None and sched.get("interval", "epoch") == "epoch":
                sched["scheduler"].step()
            for cb in self.callbacks:
                hook = getattr(cb, "on_train_epoch_end", None)
                if hook is not None:
                    hook(self, model)
            if done:
                break
        return None

    def test(self, model):
        raise NotImplementedError("the stand-in Trainer runs the training loop only")
