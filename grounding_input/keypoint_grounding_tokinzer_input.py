from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """keypoints (reference grounding_input/keypoint_grounding_tokinzer_input.py:10-44)."""
    FIELDS = (("points", "points"), ("masks", "masks"))
    ANCHOR = "points"

    def _remember(self, a):
        self.max_persons_per_image = int(a.shape[1] / 17)
