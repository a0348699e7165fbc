"""`Expand2square` as evaluation_aqa_dataset.py:28 imports it (reference minigpt4/processors/transform.py:54-107): pad a PIL
image to a square canvas, centred, and move box / point labels with it. The box-text formatters of the reference file
(:110-148) belong to the grounding datasets, not to this path."""
from PIL import Image


def expand2square(pil_img, background_color=(255, 255, 255)):
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def _shift(w, h):
    side = max(w, h)
    return (side - w) // 2, (side - h) // 2, side


def box_xyxy_expand2square(box, *, w, h):
    dx, dy, _ = _shift(w, h)
    x1, y1, x2, y2 = box
    return (x1 + dx, y1 + dy, x2 + dx, y2 + dy)


def point_xy_expand2square(point, *, w, h):
    dx, dy, _ = _shift(w, h)
    return (point[0] + dx, point[1] + dy)


class Expand2square:
    def __init__(self, background_color=(255, 255, 255)):
        self.background_color = background_color

    def __call__(self, image, labels=None):
        w, h = image.size
        out = expand2square(image, background_color=self.background_color)
        if labels is not None:
            if "boxes" in labels:
                labels["boxes"] = [box_xyxy_expand2square(b, w=w, h=h) for b in labels["boxes"]]
            if "points" in labels:
                labels["points"] = [point_xy_expand2square(p, w=w, h=h) for p in labels["points"]]
        return out, labels
