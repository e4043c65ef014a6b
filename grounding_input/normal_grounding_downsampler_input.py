class GroundingDSInput:
    """The extra input of the diffusion model = the raw map (reference grounding_input/normal_grounding_downsampler_input.py:10-16)."""

    def prepare(self, batch):
        return batch["normal"]
