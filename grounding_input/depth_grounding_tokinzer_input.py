from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """depth maps + per-sample mask (reference grounding_input/depth_grounding_tokinzer_input.py:10-43)."""
    FIELDS = (("depth", "depth"), ("mask", "mask"))
    ANCHOR = "depth"

    def _remember(self, a):
        self.C, self.H, self.W = a.shape[1:]
