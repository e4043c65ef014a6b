from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """surface-normal maps + per-sample mask (reference grounding_input/normal_grounding_tokinzer_input.py:10-43)."""
    FIELDS = (("normal", "normal"), ("mask", "mask"))
    ANCHOR = "normal"

    def _remember(self, a):
        self.C, self.H, self.W = a.shape[1:]
