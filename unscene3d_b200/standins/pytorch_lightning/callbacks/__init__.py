class Callback:
    def on_train_epoch_end(self, trainer, pl_module):
        pass


class ModelCheckpoint(Callback):
    def __init__(self, *args, **kwargs):
        self.kwargs = kwargs


class LearningRateMonitor(Callback):
    def __init__(self, *args, **kwargs):
        pass
