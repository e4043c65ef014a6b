"""Grounding tokeniser container (text_image); the MLP + Fourier features run in gligen_b200.engine
(kernels glg_position_features + glg_gemm).  Reference: ldm/modules/diffusionmodules/text_image_grounding_net.py."""
from ldm.modules.diffusionmodules.grounding_common import make_position_net


class PositionNet(make_position_net("text_image")):
    pass
