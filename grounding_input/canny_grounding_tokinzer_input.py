from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """Canny edge maps + per-sample mask (reference grounding_input/canny_grounding_tokinzer_input.py:10-43)."""
    FIELDS = (("canny_edge", "canny_edge"), ("mask", "mask"))
    ANCHOR = "canny_edge"

    def _remember(self, a):
        self.C, self.H, self.W = a.shape[1:]
