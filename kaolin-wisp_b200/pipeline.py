"""Pipeline: wisp.models.Pipeline (wisp/models/pipeline.py:14-53)."""
import torch.nn as nn


class Pipeline(nn.Module):
    def __init__(self, nef, tracer=None):
        super().__init__()
        self.nef = nef
        self.tracer = tracer

    def forward(self, *args, **kwargs):
        if self.tracer is not None:
            return self.tracer(self.nef, *args, **kwargs)
        return self.nef(*args, **kwargs)
