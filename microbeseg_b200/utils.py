"""Host-side helpers with the reference's names and semantics (src/utils/utils.py)."""
import numpy as np

# sizes the reference pads model inputs to (utils.py:137-138)
TESTED_IMG_SHAPES = (64, 128, 256, 320, 512, 768, 1024, 1280, 1408, 1600, 1920, 2048, 2240, 2560, 3200, 4096, 4480,
                     6080, 8192)


def model_input_pads(height, width):
    """[pad_y, pad_x] that zero_pad_model_input adds on the TOP / LEFT (utils.py:145-152)."""
    pads = []
    for size in (height, width):
        fit = [s for s in TESTED_IMG_SHAPES if size <= s]
        if fit:
            pads.append(fit[0] - size)
    return pads


def zero_pad_model_input(img, pad_val=0):
    """Pad the model input on the top/left up to the next tested size (utils.py:124-163).

    Returns ``(padded img, [pad_y, pad_x])``.  Mirrors the reference including its quirks: a 3-D
    input is treated as (x, y, z) and transposed around the padding, and an image larger than
    8192 in a dimension yields fewer than two pads instead of an error unless both are too big."""
    img = np.asarray(img)
    if img.ndim == 3:
        img = np.transpose(img, (2, 1, 0))
    pads = model_input_pads(img.shape[0], img.shape[1])
    if not pads:
        raise Exception('Image too big to pad. Use sliding windows')
    widths = ((pads[0], 0), (pads[1], 0)) + (((0, 0),) if img.ndim == 3 else ())
    img = np.pad(img, widths, mode='constant', constant_values=pad_val)
    if img.ndim == 3:
        img = np.transpose(img, (2, 1, 0))
    return img, [pads[0], pads[1]]


def min_max_normalization(img, min_value=None, max_value=None):
    """Clip to [min, max] and map to [-1, 1] as float32 (utils.py:50-74)."""
    if max_value is None:
        max_value = img.max()
    if min_value is None:
        min_value = img.min()
    img = np.clip(img, min_value, max_value)
    img = 2 * (img.astype(np.float32) - min_value) / (max_value - min_value) - 1
    return img.astype(np.float32)


def get_nucleus_ids(img):
    """Sorted ids > 0 of an intensity-coded label image (utils.py:11-22)."""
    values = np.unique(img)
    return values[values > 0]
