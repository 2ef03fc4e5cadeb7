"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the evaluation metric the reference applies to the masks this path produces:
  * ``label_instances``  = ``skimage.measure.label`` on an integer instance image, as called at
    /root/reference/src/evaluation/eval.py:261,313 (8-connectivity for 2-D input, components of equal non-zero value,
    numbered 1..n in raster order of their first pixel).  scikit-image is not installed here: restated with
    scipy.sparse.csgraph.connected_components over the equal-value 8-neighbour graph.  PARITY UNPINNED (no skimage).
  * ``aji_plus``         = ``get_fast_aji_plus`` (/root/reference/src/evaluation/stats_utils.py:98-179), restated with the
    same pairwise intersection / union tables, the same ``scipy.optimize.linear_sum_assignment(-iou)`` pairing and the
    same treatment of unpaired instances.  The reference module itself cannot be imported here (needs cv2).
"""
import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def label_instances(img):
    img = np.asarray(img)
    H, W = img.shape
    idx = np.arange(H * W).reshape(H, W)
    rows, cols = [], []
    for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
        ys, yd = slice(0, H - dy), slice(dy, H)
        xs, xd = (slice(0, W - dx), slice(dx, W)) if dx >= 0 else (slice(-dx, W), slice(0, W + dx))
        a, b = img[ys, xs], img[yd, xd]
        m = (a == b) & (a != 0)
        rows.append(idx[ys, xs][m])
        cols.append(idx[yd, xd][m])
    r, c = np.concatenate(rows), np.concatenate(cols)
    g = coo_matrix((np.ones(len(r), np.int8), (r, c)), shape=(H * W, H * W))
    _, comp = connected_components(g, directed=False)
    comp = comp.reshape(H, W)
    fg = img != 0
    out = np.zeros((H, W), np.int32)
    if not fg.any():
        return out
    ids, first = np.unique(comp[fg], return_index=True)          # first raster occurrence of each component
    order = np.argsort(first, kind="stable")
    lut = np.zeros(comp.max() + 1, np.int32)
    lut[ids[order]] = np.arange(1, len(ids) + 1)
    out[fg] = lut[comp[fg]]
    return out


def aji_plus(true, pred):
    """stats_utils.py:98-179; ``true`` / ``pred`` carry contiguous ids 1..n (they come from measure.label)."""
    true, pred = np.asarray(true), np.asarray(pred)
    true_ids = list(np.unique(true))
    pred_ids = list(np.unique(pred))
    if true_ids[0] != 0:
        true_ids = [0] + true_ids
    if pred_ids[0] != 0:
        pred_ids = [0] + pred_ids
    nt, npred = len(true_ids) - 1, len(pred_ids) - 1
    t_area = np.array([(true == t).sum() for t in true_ids[1:]], np.float64)
    p_area = np.array([(pred == p).sum() for p in pred_ids[1:]], np.float64)
    inter = np.zeros((nt, npred), np.float64)
    union = np.zeros((nt, npred), np.float64)
    for t in true_ids[1:]:
        overlap = np.unique(pred[true == t])
        for p in overlap:
            if p == 0:
                continue
            i = float(((true == t) & (pred == p)).sum())
            inter[t - 1, p - 1] = i
            union[t - 1, p - 1] = t_area[t - 1] + p_area[p - 1] - i
    iou = inter / (union + 1.0e-6)
    pt, pp_ = linear_sum_assignment(-iou)
    piou = iou[pt, pp_]
    pt, pp_ = pt[piou > 0.0], pp_[piou > 0.0]
    overall_inter = inter[pt, pp_].sum()
    overall_union = union[pt, pp_].sum()
    paired_t, paired_p = set((pt + 1).tolist()), set((pp_ + 1).tolist())
    for t in true_ids[1:]:
        if t not in paired_t:
            overall_union += t_area[t - 1]
    for p in pred_ids[1:]:
        if p not in paired_p:
            overall_union += p_area[p - 1]
    return overall_inter / overall_union
