"""Grounding tokeniser container for depth maps (configs/cc3m_depth.yaml); ConvNeXt-tiny + mask / position embedding + MLP run in
gligen_b200.spatial (glg_patchify_*, glg_dwconv7_ln, glg_gemm, glg_spatial_tokens).
Reference: ldm/modules/diffusionmodules/depth_grounding_net.py."""
from ldm.modules.diffusionmodules.grounding_common import make_position_net


class PositionNet(make_position_net("depth")):
    pass
