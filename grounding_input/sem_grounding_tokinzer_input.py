from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """semantic (one-hot) maps + per-sample mask (reference grounding_input/sem_grounding_tokinzer_input.py:10-43)."""
    FIELDS = (("sem", "sem"), ("mask", "mask"))
    ANCHOR = "sem"

    def _remember(self, a):
        self.C, self.H, self.W = a.shape[1:]
