"""Inpainting mask from grounding boxes (reference inpaint_mask_func.py:16-41, both random flags False):
1 = keep the known latent, 0 = inside a box; pixel coords are int-truncated box*size."""
import torch


def draw_masks_from_boxes(boxes, size, randomize_fg_mask=False, random_add_bg_mask=False):
    if randomize_fg_mask or random_add_bg_mask:
        raise NotImplementedError("random stroke masks are training-time augmentation (out of scope)")
    masks = torch.ones(boxes.shape[0], size, size)
    px = (boxes.detach().float().cpu() * size).to(torch.int64)        # trunc toward zero == int()
    for bi in range(px.shape[0]):
        for x0, y0, x1, y1 in px[bi].tolist():
            masks[bi, y0:y1, x0:x1] = 0
    return masks.unsqueeze(1)
