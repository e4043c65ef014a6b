"""Grounding input adapters: `GroundingNetInput.prepare(batch) -> kwargs` for the grounding tokenizer and
`get_null_input()` -> zero tensors of the remembered shapes (reference grounding_input/*_tokinzer_input.py).
One table-driven base class; the three modules below only name their fields."""
import torch

from gligen_b200._overlay import extend as _extend

__path__ = _extend(__path__, __name__)          # the reference's other adapters (hed / canny / depth / ...) stay importable


class TableGroundingNetInput:
    #: (kwarg name, batch key)
    FIELDS = ()
    #: kwarg whose shape/device/dtype is remembered for get_null_input
    ANCHOR = None

    def __init__(self):
        self.set = False

    def prepare(self, batch):
        self.set = True
        out = {kw: batch[key] for kw, key in self.FIELDS}
        self._shapes = {kw: tuple(v.shape[1:]) for kw, v in out.items()}
        a = out[self.ANCHOR]
        self.batch, self.device, self.dtype = a.shape[0], a.device, a.dtype
        self._remember(a)
        return out

    def _remember(self, anchor):
        pass

    def get_null_input(self, batch=None, device=None, dtype=None):
        assert self.set, "not set yet, cannot call this funcion"
        batch = self.batch if batch is None else batch
        device = self.device if device is None else device
        dtype = self.dtype if dtype is None else dtype
        return {kw: torch.zeros((batch,) + shp, device=device, dtype=dtype) for kw, shp in self._shapes.items()}
