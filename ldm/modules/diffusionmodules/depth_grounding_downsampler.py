"""Grounding downsampler container for depth maps: the planes concatenated to the latent in front of the first conv
(openaimodel.py:441-443) are produced by gligen_b200.spatial.emit_downsampler (glg_resize_plane, glg_conv2d_small).
Reference: ldm/modules/diffusionmodules/depth_grounding_downsampler.py."""
from ldm.modules.diffusionmodules.grounding_common import make_downsampler


class GroundingDownsampler(make_downsampler("depth")):
    pass
