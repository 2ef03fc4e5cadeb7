"""Host-side mirror of the numerical part of the reference's analysis export, on libmbseg (CUDA).

``frame_statistics(masks)`` reproduces the per-frame table of /root/reference/src/inference/analysis.py:141-170 for a
[T,H,W] (or [H,W]) instance mask stack: ``frame``, ``counts`` (= max id, :151), ``total_area`` (= the SUM OF THE LABEL
VALUES, a quirk of :152 that is kept), ``mean_area``, ``mean_minor_axis_length``, ``mean_major_axis_length`` (means over
``regionprops`` in ascending id order, :154-166).  Areas and second moments of all instances of all frames come from
one pass of the label-statistics kernels (the ones label generation uses); no CPU fallback.
"""
import numpy as np
import torch

from . import _native as nat


def instance_statistics(masks):
    """(area int32 [T, ids], major float64 [T, ids], minor float64 [T, ids]) with ids = max id + 1 (row 0 = background)."""
    m = np.ascontiguousarray(masks)
    if m.ndim == 2:
        m = m[None]
    if m.dtype != np.uint16:
        if m.min() < 0 or m.max() > 65535:
            raise ValueError("instance ids must fit in uint16")
        m = m.astype(np.uint16)
    if not torch.cuda.is_available():
        raise RuntimeError("microbeseg_b200.analysis needs a CUDA device (no CPU fallback)")
    L = nat.lib()
    T, H, W = m.shape
    max_id = int(m.max())
    ids = max_id + 1
    dev = torch.from_numpy(m.view(np.int16)).cuda()
    area = torch.empty((T, ids), dtype=torch.int32, device=dev.device)
    major = torch.empty((T, ids), dtype=torch.float64, device=dev.device)
    minor = torch.empty((T, ids), dtype=torch.float64, device=dev.device)
    ws = torch.empty(T * ids * 96 + 4096, dtype=torch.uint8, device=dev.device)
    with torch.cuda.device(dev.device):
        nat.check(L.mbs_instance_stats(dev.data_ptr(), T, H, W, max_id, area.data_ptr(), major.data_ptr(), minor.data_ptr(),
                                       ws.data_ptr(), ws.numel(), nat.stream_ptr()), "instance_stats")
    return area.cpu().numpy(), major.cpu().numpy(), minor.cpu().numpy()


def frame_statistics(masks):
    """dict of lists, one entry per frame, with the columns of analysis.py:143-166."""
    m = np.asarray(masks)
    if m.ndim == 2:
        m = m[None]
    area, major, minor = instance_statistics(m)
    res = {'frame': [], 'counts': [], 'mean_area': [], 'total_area': [], 'mean_minor_axis_length': [],
           'mean_major_axis_length': []}
    for t in range(len(m)):
        present = area[t, 1:] > 0                    # regionprops lists the ids that occur, ascending
        res['frame'].append(t)
        res['counts'].append(m[t].max())
        res['total_area'].append(np.sum(m[t]))
        res['mean_area'].append(np.mean(area[t, 1:][present].astype(np.float64)) if present.any() else np.nan)
        res['mean_minor_axis_length'].append(np.mean(minor[t, 1:][present]) if present.any() else np.nan)
        res['mean_major_axis_length'].append(np.mean(major[t, 1:][present]) if present.any() else np.nan)
    return res
