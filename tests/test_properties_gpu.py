"""GPU: size-independent properties of the denoiser at BASELINE's full model size (SD-1.4 GLIGEN box+text, G = 30), through the drop-in
UNetModel - checks that need no reference output:
  * determinism: the same inputs give bit-identical eps (every kernel is deterministic: fixed-order reductions, no atomics on values);
  * row independence: a sample's eps does not change when ANOTHER sample of the batch changes (GroupNorm / LayerNorm / attention are
    per sample; same batch size -> same tiles and summation order -> bit-identical);
  * the grounding objects are a SET: permuting the object slots (boxes, masks, embeddings together) changes eps only through fp32
    summation order inside the fuser attention - bounded by the per-forward tolerance of DESIGN 2 (a slot-order dependence would be O(1));
  * null grounding == all masks zero: `grounding_input` absent (get_null_input) equals passing zero masks explicitly."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def case():
    from gligen_b200 import synth
    from gligen_b200.pipeline import build_model, to_device
    cfg, model = build_model("sd14_box_text", device=DEV)
    inp = synth.make_inputs(cfg, 2, 30, seed=17, n_valid=7)
    batch = to_device(inp["batch"], DEV)
    x, ctx = inp["x"].to(DEV), inp["context"].to(DEV)
    ts = torch.tensor([801, 321], device=DEV)
    return cfg, model, batch, x, ctx, ts


def _eps(model, batch, x, ctx, ts):
    g = model.grounding_tokenizer_input.prepare(batch)
    out = model(dict(x=x, timesteps=ts, context=ctx, grounding_input=g, inpainting_extra_input=None, grounding_extra_input=None)).clone()
    torch.cuda.synchronize()
    return out


def test_deterministic(case):
    cfg, model, batch, x, ctx, ts = case
    a = _eps(model, batch, x, ctx, ts)
    model.invalidate_static()
    b = _eps(model, {k: v.clone() for k, v in batch.items()}, x.clone(), ctx.clone(), ts)
    assert torch.equal(a, b)


def test_rows_are_independent(case):
    cfg, model, batch, x, ctx, ts = case
    a = _eps(model, batch, x, ctx, ts)
    x2, ctx2 = x.clone(), ctx.clone()
    x2[1] = torch.randn_like(x2[1]); ctx2[1] = torch.randn_like(ctx2[1])
    b2 = {k: v.clone() for k, v in batch.items()}
    b2["boxes"][1] = b2["boxes"][1].flip(0); b2["text_embeddings"][1] = -b2["text_embeddings"][1]
    b = _eps(model, b2, x2, ctx2, ts)
    assert torch.equal(a[0], b[0]), f"row 0 moved by {rel_l2(b[0], a[0]):.3e} when only row 1 changed"
    assert rel_l2(b[1], a[1]) > 1e-2


def test_grounding_objects_are_a_set(case):
    cfg, model, batch, x, ctx, ts = case
    a = _eps(model, batch, x, ctx, ts)
    perm = torch.randperm(batch["boxes"].shape[1], generator=torch.Generator().manual_seed(3)).to(DEV)
    b2 = {k: v[:, perm].contiguous() for k, v in batch.items()}
    b = _eps(model, b2, x, ctx, ts)
    r = rel_l2(b, a)
    print(f"\\nobject-slot permutation: eps rel-L2 {r:.3e}")
    assert r <= 4e-2          # ulp-level changes are amplified to ~1.5e-2 by the random-weight UNet; a slot-order bug would be O(1)


def test_null_grounding_is_zero_masks(case):
    cfg, model, batch, x, ctx, ts = case
    g = model.grounding_tokenizer_input.prepare(batch)
    null = model(dict(x=x, timesteps=ts, context=ctx, inpainting_extra_input=None, grounding_extra_input=None)).clone()
    zero = {k: torch.zeros_like(v) for k, v in g.items()}
    explicit = model(dict(x=x, timesteps=ts, context=ctx, grounding_input=zero, inpainting_extra_input=None, grounding_extra_input=None)).clone()
    torch.cuda.synchronize()
    assert torch.equal(null, explicit)
