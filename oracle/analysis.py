"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Per-frame analysis table of /root/reference/src/inference/analysis.py:141-170 (counts, total_area = sum of the label
values (sic, :152), mean area / axis lengths over regionprops).  ``regionprops`` is the restatement in
oracle/labels.py (PARITY UNPINNED: scikit-image is not installable here)."""
import numpy as np

from . import labels as ol


def frame_statistics(masks):
    m = np.asarray(masks)
    if m.ndim == 2:
        m = m[None]
    res = {'frame': [], 'counts': [], 'mean_area': [], 'total_area': [], 'mean_minor_axis_length': [],
           'mean_major_axis_length': []}
    for t in range(len(m)):
        res['frame'].append(t)
        res['counts'].append(np.max(m[t]))
        res['total_area'].append(np.sum(m[t]))
        props = ol.regionprops(m[t])
        areas = [p.area for p in props]
        res['mean_area'].append(np.mean(np.array(areas)) if areas else np.nan)
        res['mean_minor_axis_length'].append(np.mean(np.array([p.minor_axis_length for p in props])) if areas else np.nan)
        res['mean_major_axis_length'].append(np.mean(np.array([p.major_axis_length for p in props])) if areas else np.nan)
    return res
