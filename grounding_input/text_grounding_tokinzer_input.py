from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """box + text (reference grounding_input/text_grounding_tokinzer_input.py:10-45)."""
    FIELDS = (("boxes", "boxes"), ("masks", "masks"), ("positive_embeddings", "text_embeddings"))
    ANCHOR = "positive_embeddings"

    def _remember(self, a):
        self.max_box, self.in_dim = a.shape[1], a.shape[2]
