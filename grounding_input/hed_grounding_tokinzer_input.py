from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """HED edge maps + per-sample mask (reference grounding_input/hed_grounding_tokinzer_input.py:10-43)."""
    FIELDS = (("hed_edge", "hed_edge"), ("mask", "mask"))
    ANCHOR = "hed_edge"

    def _remember(self, a):
        self.C, self.H, self.W = a.shape[1:]
