"""Natural-synthetic-anomaly (NSA) cut-paste augmentation that builds the `aug_image` half of every training batch
(reference minigpt4/datasets/self_sup_tasks.py:11-292, adopted there from AnomalyGPT / NSA; called from
anomaly_detection.py:244-262 with the per-class bounds of :44-62).

`patch_ex(dest, src, ...) -> (augmented uint8 image, label map [H, W, 1], boxes [[x0, y0, x1, y1], ...])`:
up to `num_patches` rectangles are cut from `src` (half-widths ~ lower + Gamma(shape, scale) of the image side, clipped to
`width_bounds_pct`), optionally rescaled (N(1, 0.5) clipped to `resize_bounds`), moved to a random place of `dest` and
blended in (Poisson cloning by default); with `skip_background` a patch must cover object pixels in the source and overlap
the destination object. The label is the median-filtered mean absolute intensity change inside the pasted area, squashed by
a logistic (`logistic-intensity`), thresholded at `tol` (`binary`), or the raw filtered change (`intensity`).

Written from the published algorithm for the options the Myriad datasets use; `num_ellipses` and
`cutpaste_patch_generation` (unused by them) are rejected. Host-side numpy / OpenCV: this is data preparation, not the
GPU hot path."""
import cv2
import numpy as np
from scipy import ndimage

_MAX_TRIES = 200


def _disk(radius):
    y, x = np.ogrid[-radius:radius + 1, -radius:radius + 1]
    return (x * x + y * y) <= radius * radius


def _object_mask(img, specs):
    """1 where the pixel's mean intensity is further than `threshold` from every listed background level."""
    m = np.ones(img.shape[:2], dtype=np.uint8)
    grey = img.mean(axis=-1)
    for level, threshold in specs:
        m &= (np.abs(grey - level) > threshold).astype(np.uint8)
    return cv2.medianBlur(m, 7)[..., None]


def _half_widths(shape, bounds, gamma_params):
    lo = [int(round(bounds[d][0] * shape[d])) for d in (0, 1)]
    hi = [int(round(bounds[d][1] * shape[d])) for d in (0, 1)]
    if gamma_params is not None:
        k, theta, offset = gamma_params
        hw = [int(np.clip((offset + np.random.gamma(k, theta)) * shape[d], lo[d], hi[d])) for d in (0, 1)]
    else:
        hw = [np.random.randint(lo[d], hi[d]) for d in (0, 1)]
    return lo, hi, hw


def _paste_one(dest, src, dest_obj, src_obj, mode, factor, shift, resize, bounds, gamma_params, min_object_pct, min_overlap_pct,
               resize_bounds, verbose):
    """-> (image, (y0, y1, x0, x1), patch mask) or (dest copy, None, None) when no admissible patch / place was found."""
    H, W = dest.shape[:2]
    lo, hi, hw = _half_widths(dest.shape, bounds, gamma_params)
    use_obj = src_obj is not None and dest_obj is not None
    for attempt in range(_MAX_TRIES):  # source rectangle
        cy, cx = np.random.randint(lo[0], H - lo[0]), np.random.randint(lo[1], W - lo[1])
        y0, y1 = max(cy - hw[0], 0), min(cy + hw[0], H)
        x0, x1 = max(cx - hw[1], 0), min(cx + hw[1], W)
        if not use_obj or src_obj[y0:y1, x0:x1].sum() / float((y1 - y0) * (x1 - x0)) > min_object_pct:
            break
    else:
        if verbose:
            print("No suitable patch found.")
        return dest.copy(), None, None
    patch = src[y0:y1, x0:x1]
    obj = src_obj[y0:y1, x0:x1, 0] if use_obj else None
    h, w = patch.shape[:2]
    if resize:
        s = np.clip(np.random.normal(1, 0.5), *resize_bounds)
        nh = np.clip(s * h, lo[0], hi[0])
        nw = int(np.clip(int(nh / h * w), lo[1], hi[1]))
        nh = int(np.clip(int(nw / w * h), lo[0], hi[0]))
        patch = cv2.resize(patch, (nw, nh))
        if patch.ndim == 2:
            patch = patch[..., None]
        if use_obj:
            obj = cv2.resize(obj, (nw, nh))
        h, w = nh, nw
    pmask = np.ones((h, w, 1), dtype=np.uint8)
    if use_obj:
        obj = obj[..., None]
    if shift:
        for attempt in range(_MAX_TRIES):  # destination centre
            cy, cx = np.random.randint(h // 2 + 1, H - h // 2 - 1), np.random.randint(w // 2 + 1, W - w // 2 - 1)
            y0, y1, x0, x1 = cy - h // 2, cy + (h + 1) // 2, cx - w // 2, cx + (w + 1) // 2
            if not use_obj:
                break
            d = dest_obj[y0:y1, x0:x1]
            n_obj = float(obj.sum())
            if n_obj / (h * w) > min_object_pct and (d & obj & pmask).sum() / max(n_obj, 1.0) > min_overlap_pct:
                break
        else:
            if verbose:
                print("No suitable center found. Dims were:", w, h)
            return dest.copy(), None, None
    else:
        y1, x1 = y0 + h, x0 + w
        if y1 > H or x1 > W:
            return dest.copy(), None, None
    if use_obj:
        pmask = pmask & (obj | dest_obj[y0:y1, x0:x1])
    out = dest.copy()
    region = out[y0:y1, x0:x1]
    if mode == "swap":
        region[...] = np.where(pmask > 0, patch, region)
    elif mode == "uniform":
        region[...] = np.uint8(np.floor(region + factor * pmask * (patch.astype(np.float64) - region)))
    elif mode in (cv2.NORMAL_CLONE, cv2.MIXED_CLONE):
        clone = pmask
        if use_obj:  # pure background on both sides joins the mask so the cloning boundary stays artefact free
            clone = pmask | ((1 - obj) & (1 - dest_obj[y0:y1, x0:x1]))
        clone = (np.uint8(np.ceil(factor * 255)) * clone).astype(np.uint8)
        clone[0], clone[-1], clone[:, 0], clone[:, -1] = 0, 0, 0, 0
        if (clone > 0).sum() < 50:  # seamlessClone fails on tiny masks
            return dest.copy(), None, None
        centre = (x1 - (x1 - x0) // 2, y0 + (y1 - y0) // 2)
        try:
            if dest.shape[2] == 1:
                z = np.zeros_like
                out = cv2.seamlessClone(np.concatenate([patch, z(patch), z(patch)], 2), np.concatenate([dest, z(dest), z(dest)], 2),
                                        clone, centre, mode)[..., :1]
            else:
                out = cv2.seamlessClone(np.ascontiguousarray(patch), dest, clone, centre, mode)
        except cv2.error as e:
            print("WARNING, tried bad interpolation mask and got:", e)
            return dest.copy(), None, None
    else:
        raise ValueError("mode not supported" + str(mode))
    return out, (y0, y1, x0, x1), pmask


def patch_ex(ima_dest, ima_src=None, same=False, num_patches=1, mode=cv2.NORMAL_CLONE, width_bounds_pct=((0.05, 0.2), (0.05, 0.2)),
             min_object_pct=0.25, min_overlap_pct=0.25, shift=True, label_mode="binary", skip_background=None, tol=1, resize=True,
             gamma_params=None, intensity_logistic_params=(1 / 6, 20), resize_bounds=(0.7, 1.3), num_ellipses=None, verbose=True,
             cutpaste_patch_generation=False):
    if num_ellipses is not None or cutpaste_patch_generation:
        raise NotImplementedError("num_ellipses / cutpaste_patch_generation are not used by the Myriad datasets")
    if mode == "mix":
        mode = (cv2.NORMAL_CLONE, cv2.MIXED_CLONE)[np.random.randint(2)]
    ima_src = ima_dest.copy() if (same or ima_src is None) else ima_src
    src_obj = dest_obj = None
    if skip_background is not None:
        specs = [skip_background] if isinstance(skip_background, tuple) else list(skip_background)
        src_obj, dest_obj = _object_mask(ima_src, specs), _object_mask(ima_dest, specs)
    factor = np.random.uniform(0.05, 0.95) if label_mode == "continuous" else 1
    H, W = ima_dest.shape[:2]
    mask = np.zeros((H, W, 1), dtype=ima_dest.dtype)
    out = ima_dest.copy()
    hull = [H - 1, 0, W - 1, 0]  # running bounding hull (y0, y1, x0, x1) of every pasted patch, as the reference reports it
    boxes = []
    for i in range(num_patches):
        if i > 0 and np.random.randint(2) == 0:  # the first patch is always attempted, every further one on a coin flip
            continue
        out, where, pmask = _paste_one(out, ima_src, dest_obj, src_obj, mode, factor, shift, resize, width_bounds_pct, gamma_params,
                                       min_object_pct, min_overlap_pct, resize_bounds, verbose)
        if pmask is None:
            continue
        y0, y1, x0, x1 = where
        mask[y0:y1, x0:x1] = pmask
        hull = [min(hull[0], y0), max(hull[1], y1), min(hull[2], x0), max(hull[3], x1)]
        boxes.append([hull[2], hull[0], hull[3], hull[1]])
    diff = np.mean(np.abs(mask * (ima_dest * 1.0) - mask * (out * 1.0)), axis=-1, keepdims=True)
    changed = np.uint8(diff > tol)
    changed[..., 0] = cv2.medianBlur(changed[..., 0], 5)
    if label_mode == "binary":
        label = changed
    elif label_mode == "continuous":
        label = changed * factor
    elif label_mode in ("intensity", "logistic-intensity"):
        label = np.mean(np.abs(changed * (ima_dest * 1.0) - changed * (out * 1.0)), axis=-1, keepdims=True)
        label[..., 0] = ndimage.median_filter(label[..., 0], footprint=_disk(5), mode="nearest")
        if label_mode == "logistic-intensity":
            k, x0 = intensity_logistic_params
            label = changed / (1 + np.exp(-k * (label - x0)))
    else:
        raise ValueError("label_mode not supported" + str(label_mode))
    return out, label, boxes
