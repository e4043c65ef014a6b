class GroundingDSInput:
    """The extra input of the diffusion model = the raw map (reference grounding_input/sem_grounding_downsampler_input.py:10-16)."""

    def prepare(self, batch):
        return batch["sem"]
