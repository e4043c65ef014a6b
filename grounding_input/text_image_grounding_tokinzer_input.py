from . import TableGroundingNetInput


class GroundingNetInput(TableGroundingNetInput):
    """box + text + image (reference grounding_input/text_image_grounding_tokinzer_input.py:10-63)."""
    FIELDS = (("boxes", "boxes"), ("masks", "masks"), ("text_masks", "text_masks"), ("image_masks", "image_masks"),
              ("text_embeddings", "text_embeddings"), ("image_embeddings", "image_embeddings"))
    ANCHOR = "text_embeddings"

    def _remember(self, a):
        self.max_box, self.in_dim = a.shape[1], a.shape[2]
