class _Logger:
    def __init__(self, *args, **kwargs):
        self.kwargs = kwargs
        self.hyperparams = None

    def log_hyperparams(self, params, *args, **kwargs):
        self.hyperparams = params

    def log_metrics(self, *args, **kwargs):
        pass


class WandbLogger(_Logger):
    pass


class CSVLogger(_Logger):
    pass


class TensorBoardLogger(_Logger):
    pass
