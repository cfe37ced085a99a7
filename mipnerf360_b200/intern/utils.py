"""intern/utils.py:17-21 — host helper used by model.render_image."""
import numpy as np


def to8b(img):
    """intern/utils.py:17-21: clip to [0,1], scale to uint8."""
    if len(img.shape) >= 3:
        return np.array([to8b(i) for i in img])
    return (255 * np.clip(np.nan_to_num(img), 0, 1)).astype(np.uint8)
