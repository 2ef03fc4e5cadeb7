"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's training-label generation for the distance method:

  * ``get_label(mask, 'distance', max_mal)``   -- src/training/train_data_representations.py:11-37
  * ``distance_label(label, search_radius)``   -- train_data_representations.py:261-361
  * ``border_label``                           -- train_data_representations.py:102-126
  * ``bottom_hat_closing``                     -- train_data_representations.py:40-72
  * ``max_mal`` of CreateLabelsWorker          -- src/training/train.py:74-79

The reference module cannot be imported here (it needs scikit-image 0.19 and cv2 at import time and
uses ``np.float``, removed in NumPy >= 1.24).  Everything that is scipy in the reference stays a
real scipy call (distance_transform_edt, binary_dilation/erosion/closing, grey_closing,
generate_binary_structure); the scikit-image pieces are restated:

  * ``measure.regionprops``: ``label``, ``area``, ``centroid`` (mean of pixel coordinates, float64),
    ``major/minor_axis_length`` = 4*sqrt(eigenvalue of the inertia tensor of the central second
    moments / area) (skimage/measure/_regionprops.py, _moments.py::inertia_tensor);
  * ``measure.label`` on a 2-D uint8 image: full (8-) connectivity, raster-first numbering;
  * ``morphology.disk(3)``: 7x7 mask x^2+y^2 <= 9.

PARITY UNPINNED for the scikit-image pieces (no reference tests / fixtures exist, scikit-image is
not installable here); the scipy pieces are the real thing.
"""
import numpy as np
from scipy import ndimage
from scipy.ndimage import distance_transform_edt, generate_binary_structure, grey_closing


def disk(radius):
    r = np.arange(-radius, radius + 1)
    return ((r[:, None] ** 2 + r[None, :] ** 2) <= radius * radius).astype(np.uint8)


def get_nucleus_ids(img):
    v = np.unique(img)
    return v[v > 0]


def label8(binary):
    return ndimage.label(binary, structure=np.ones((3, 3), bool))


class Region:
    """The regionprops attributes the reference reads."""

    def __init__(self, label_id, coords):
        self.label = int(label_id)
        self.coords = coords                      # (n, 2) int64 rows, cols
        self.area = coords.shape[0]
        self.centroid = tuple(coords.mean(axis=0))
        self._axes = None

    def _axis_lengths(self):
        if self._axes is None:
            d = self.coords.astype(np.float64) - np.asarray(self.centroid)
            mu20 = float((d[:, 0] ** 2).sum())
            mu02 = float((d[:, 1] ** 2).sum())
            mu11 = float((d[:, 0] * d[:, 1]).sum())
            n = float(self.area)
            t = np.array([[mu02 / n, -mu11 / n], [-mu11 / n, mu20 / n]])
            ev = np.clip(np.linalg.eigvalsh(t), 0, None)
            self._axes = (4 * np.sqrt(ev.max()), 4 * np.sqrt(ev.min()))
        return self._axes

    @property
    def major_axis_length(self):
        return self._axes[0] if self._axes else self._axis_lengths()[0]

    @property
    def minor_axis_length(self):
        return self._axes[1] if self._axes else self._axis_lengths()[1]


def regionprops(label):
    label = np.asarray(label)
    out = []
    ids = get_nucleus_ids(label)
    if len(ids) == 0:
        return out
    ys, xs = np.nonzero(label)
    vals = label[ys, xs]
    order = np.argsort(vals, kind="stable")        # raster order inside each id is preserved
    ys, xs, vals = ys[order], xs[order], vals[order]
    bounds = np.flatnonzero(np.diff(vals)) + 1
    for seg_y, seg_x, seg_v in zip(np.split(ys, bounds), np.split(xs, bounds), np.split(vals, bounds)):
        out.append(Region(seg_v[0], np.stack([seg_y, seg_x], 1).astype(np.int64)))
    return out


def max_major_axis_length(mask):
    """train.py:74-79 -> int(ceil(max major_axis_length))"""
    props = regionprops(mask)
    return int(np.ceil(np.max(np.array([p.major_axis_length for p in props]))))


def boundary_label(label):
    """train_data_representations.py:75-99"""
    label_bin = label > 0
    kernel = np.ones((3, 3), np.uint8)
    boundary = np.zeros(label.shape, bool)
    for nucleus_id in get_nucleus_ids(label):
        nucleus = label == nucleus_id
        boundary |= ndimage.binary_dilation(nucleus, kernel) ^ nucleus
    return np.maximum(label_bin, 2 * boundary).astype(np.uint8)


def cell_distance_label(label, search_radius, apply_clipping=False, clip_val=5):
    """train_data_representations.py:220-258"""
    label_dist = np.zeros(label.shape, np.float64)
    for prop in regionprops(label):
        nucleus = label == prop.label
        ys, xs = _window(prop.centroid, search_radius, label.shape)
        d = distance_transform_edt(nucleus[ys, xs])
        if d.max() > 0 and not apply_clipping:
            d = d / d.max()
        label_dist[ys, xs] += d
    if apply_clipping:
        label_dist = np.clip(label_dist, 0, clip_val)      # :252-256
        label_dist = label_dist / clip_val
    return label_dist.astype(np.float32)


def border_label(label):
    """train_data_representations.py:102-126"""
    label_bin = label > 0
    kernel = np.ones((3, 3), np.uint8)
    boundary = np.zeros(label.shape, bool)
    for nucleus_id in get_nucleus_ids(label):
        nucleus = label == nucleus_id
        boundary |= ndimage.binary_dilation(nucleus, kernel) ^ nucleus
    border = boundary ^ (ndimage.binary_dilation(label_bin, kernel) ^ label_bin)
    return np.maximum(label_bin, 2 * border).astype(np.uint8)


def j4_label(label, k_neighbors=2, se_radius=4):
    """train_data_representations.py:158-190 with compute_neighbor_instances (:193-217): scipy's binary_closing (the real
    dependency) + the window count of distinct positive ids on the zero-padded mask"""
    label_bin = label > 0
    label_bottom_hat = ndimage.binary_closing(label_bin, disk(se_radius)) ^ label_bin
    k = k_neighbors
    padded = np.pad(label, pad_width=k, constant_values=0)
    n_neighbors = np.zeros(label.shape, np.int64)
    for y in range(label.shape[0]):
        for x in range(label.shape[1]):
            crop = padded[y:y + 2 * k + 1, x:x + 2 * k + 1]
            n_neighbors[y, x] = len(set(crop[crop > 0].tolist()))
    label_bg = (~label_bin) & (~label_bottom_hat)
    label_gap = (~label_bin) & label_bottom_hat
    label_touching = label_bin & (n_neighbors > 1)
    label_cell = ~(label_bg | label_gap | label_touching)
    j4 = np.maximum(label_bg, 2 * label_cell)
    j4 = np.maximum(j4, 3 * label_touching)
    j4 = np.maximum(j4, 4 * label_gap)
    j4 -= 1
    return j4.astype(np.uint8)


def bottom_hat_closing(label):
    """train_data_representations.py:40-72 -> (gap labels int, gap map float32)"""
    label_bin = np.zeros_like(label, dtype=bool)
    se = disk(3)
    for nucleus_id in get_nucleus_ids(label):
        nucleus = ndimage.binary_closing(label == nucleus_id, se)
        label_bin[nucleus] = True
    label_bottom_hat = ndimage.binary_closing(label_bin, se) ^ label_bin
    label_closed = (~label_bin) & label_bottom_hat
    label_closed, _ = label8(label_closed)
    props = regionprops(label_closed)
    corr = (label_closed > 0).astype(np.float32)
    cross = generate_binary_structure(2, 1)
    for p in props:
        if p.minor_axis_length >= 3:
            gap = label_closed == p.label
            ring = gap ^ ndimage.binary_erosion(gap, cross)
            corr[gap] = 1
            corr[ring] = 0.8
    return label_closed, corr


def _window(centroid, search_radius, shape):
    c = np.round(centroid)                                  # round half to even, as np.round
    y0, y1 = int(max(c[0] - search_radius, 0)), int(min(c[0] + search_radius, shape[0]))
    x0, x1 = int(max(c[1] - search_radius, 0)), int(min(c[1] + search_radius, shape[1]))
    return slice(y0, y1), slice(x0, x1)


def distance_label(label, search_radius, return_intermediates=False):
    """train_data_representations.py:261-361 -> (cell_dist float32, neighbor_dist float32)"""
    label = np.asarray(label)
    label_dist = np.zeros(label.shape, np.float64)
    label_dist_neighbor = np.zeros(label.shape, np.float64)
    label_border = border_label(label) == 2                                           # :276
    for p in regionprops(label):                                                      # :279-330
        win = _window(p.centroid, search_radius, label.shape)
        crop = label[win]
        nucleus_crop_dist = distance_transform_edt(crop == p.label)                   # :289
        max_dist = np.max(nucleus_crop_dist) if nucleus_crop_dist.size else 0
        if max_dist > 0:
            nucleus_crop_dist = nucleus_crop_dist / max_dist
        else:
            continue
        label_dist[win] += nucleus_crop_dist
        nb = np.copy(crop)
        if len(get_nucleus_ids(nb)) <= 1:                                             # :309
            continue
        own = nb == p.label
        nb[nb == 0] = p.label
        nb[nb != p.label] = 0
        nb_dist = distance_transform_edt(nb > 0) * own                                # :317-318
        if np.max(nb_dist) > 0:
            denominator = np.minimum(max_dist + 3, np.max(nb_dist))
            nb_dist = np.clip(nb_dist / denominator, 0, 1)
        else:
            nb_dist = 1
        label_dist_neighbor[win] += (1 - nb_dist) * own
    label_closed, label_closed_corr = bottom_hat_closing(label)                       # :333
    kernel = np.ones((3, 3), np.uint8)
    for g in regionprops(label_closed):                                               # :337-350
        obj = label_closed == g.label
        obj_boundary = ndimage.binary_dilation(obj, kernel) ^ obj
        th = 5 if g.area <= 20 else 8 if g.area <= 30 else 10 if g.area <= 50 else 20
        if np.sum(obj_boundary * label_dist_neighbor) < th:
            label_closed_corr[obj] = 0
    pre = label_dist_neighbor.copy()
    label_dist_neighbor = np.maximum(label_dist_neighbor, label_closed_corr.astype(np.float64))
    label_dist_neighbor = np.maximum(label_dist_neighbor, label_border.astype(np.float64))
    label_dist_neighbor = 1 / np.sqrt(0.65 + 0.5 * np.exp(-11 * (label_dist_neighbor - 0.75))) - 0.19   # :357
    label_dist_neighbor = np.clip(label_dist_neighbor, 0, 1)
    label_dist_neighbor = grey_closing(label_dist_neighbor, size=(3, 3))
    out = label_dist.astype(np.float32), label_dist_neighbor.astype(np.float32)
    if return_intermediates:
        return out, dict(label_border=label_border, gaps=label_closed, gap_map=label_closed_corr, neighbor_raw=pre)
    return out


def get_label(mask, label_type, max_mal):
    """train_data_representations.py:11-37 (the label types the CUDA path builds)"""
    if label_type == 'boundary':
        return boundary_label(mask)
    if label_type == 'border':
        return border_label(mask)
    if label_type == 'cell_dist':
        return cell_distance_label(mask, search_radius=int(np.ceil(0.75 * max_mal)))
    if label_type == 'cell_dist_clipped':
        return cell_distance_label(mask, search_radius=int(np.ceil(0.75 * max_mal)), apply_clipping=True)
    if label_type == 'j4':
        return j4_label(mask)
    if label_type != 'distance':
        raise Exception('Label type not known')
    return distance_label(mask, search_radius=int(np.ceil(0.75 * max_mal)))


def create_labels(mask):
    """CreateLabelsWorker.create_labels for one mask (train.py:72-84): max_mal, then the two maps."""
    max_mal = max_major_axis_length(mask)
    return get_label(mask, 'distance', max_mal), max_mal
