"""Module *types* of the transformer blocks, kept at the reference's import path.

In this engine these classes are parameter containers only: the arithmetic of
reference ldm/modules/attention.py (SelfAttention :154-186, CrossAttention :102-149,
GatedSelfAttentionDense :215-244, BasicTransformerBlock :303-338, SpatialTransformer :341-376) runs as
fused sm_100a kernels driven by gligen_b200.engine; the containers exist so that
  * state_dict() keys equal the reference's (checkpoints load verbatim), and
  * `set_alpha_scale` (gligen_inference.py:24-28), which walks model.modules() and tests
    `type(module) == GatedSelfAttentionDense`, finds the fusers and sets `.scale` on them.
"""
import torch.nn as nn


class ParamNode(nn.Module):
    """A named bag of parameters / child nodes; children with integer names are indexable like
    nn.ModuleList / nn.Sequential (`model.input_blocks[0][0]`)."""

    def __getitem__(self, idx):
        return self._modules[str(idx)]

    def __setitem__(self, idx, module):
        self._modules[str(idx)] = module

    def __len__(self):
        return len(self._modules)

    def forward(self, *args, **kwargs):
        raise RuntimeError(f"{type(self).__name__} is a parameter container; call UNetModel.forward(input) "
                           "(the whole denoiser runs inside the gligen_b200 engine)")


class SelfAttention(ParamNode):
    pass


class CrossAttention(ParamNode):
    pass


class FeedForward(ParamNode):
    pass


class GatedSelfAttentionDense(ParamNode):
    """Fuser of visual and grounding tokens.  `scale` multiplies tanh(alpha) of both gated residuals; the
    engine reads it at the next forward and skips the fuser entirely when it is 0."""

    def __init__(self):
        super().__init__()
        self.scale = 1


class GatedCrossAttentionDense(ParamNode):
    """Importable for `set_alpha_scale`; no shipped config selects fuser_type=gatedCA (not accelerated)."""

    def __init__(self):
        super().__init__()
        self.scale = 1


class GatedSelfAttentionDense2(ParamNode):
    """Importable name only; no shipped config selects fuser_type=gatedSA2 (not accelerated)."""

    def __init__(self):
        super().__init__()
        self.scale = 1


class BasicTransformerBlock(ParamNode):
    pass


class SpatialTransformer(ParamNode):
    pass
