"""Host-side mirror of the reference's mask -> polygon step, running on libmbseg (CUDA).

``mask_to_polygons(mask)`` returns, for every non-zero id of a uint16 instance mask (ids ascending, like the
groupby of get_indices_pandas, /root/reference/src/utils/hull_polygon.py:8-42), what ``cv2_countour`` (:45-89) returns
for it: a list with one (2, N) integer array, rows = (y, x), points in cv2.findContours(RETR_TREE,
CHAIN_APPROX_NONE) order.  ``points_string`` formats one polygon the way infer.py:281-284 does ("x,y x,y ...").
The per-cell cv2.findContours loop is the next CPU bottleneck once a 2048^2 frame is segmented in 10 ms; here all
instances of a frame are traced in two kernel launches.  No CPU fallback.
"""
import numpy as np
import torch

from . import _native as nat


def contours_device(mask_dev, n_labels=None):
    """``mask_dev``: (H,W) CUDA tensor, uint16 payload (torch.int16 / torch.uint16).  Returns (offsets int64 [n+1] on
    the host, points int32 [total, 2] (y, x) on the device); instance id l+1 owns points[offsets[l]:offsets[l+1]]."""
    if not mask_dev.is_cuda or mask_dev.dim() != 2:
        raise RuntimeError("polygons.contours_device needs a 2-D CUDA mask (no CPU fallback)")
    L = nat.lib()
    mask_dev = mask_dev.contiguous()
    H, W = mask_dev.shape
    dev = mask_dev.device
    with torch.cuda.device(dev):
        if n_labels is None:
            n_labels = int(mask_dev.view(torch.int16).to(torch.int32).bitwise_and(0xFFFF).max().item())
        n_labels = int(n_labels)
        if n_labels == 0:
            return np.zeros(1, np.int64), torch.zeros((0, 2), dtype=torch.int32, device=dev)
        first = torch.empty(n_labels, dtype=torch.int32, device=dev)
        counts = torch.empty(n_labels, dtype=torch.int32, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        sp = nat.stream_ptr()
        nat.check(L.mbs_contour_first(mask_dev.data_ptr(), H, W, n_labels, first.data_ptr(), sp), "contour_first")
        nat.check(L.mbs_contour_trace(mask_dev.data_ptr(), H, W, n_labels, first.data_ptr(), None, counts.data_ptr(), None,
                                      overflow.data_ptr(), sp), "contour_trace(count)")
        offsets = torch.zeros(n_labels + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=offsets[1:])
        total = int(offsets[-1].item())
        points = torch.empty((max(total, 1), 2), dtype=torch.int32, device=dev)
        nat.check(L.mbs_contour_trace(mask_dev.data_ptr(), H, W, n_labels, first.data_ptr(), offsets.data_ptr(), counts.data_ptr(),
                                      points.data_ptr(), overflow.data_ptr(), sp), "contour_trace(write)")
        if int(overflow.item()):
            raise RuntimeError("contour walk exceeded its step guard (corrupt mask?)")
    return offsets.cpu().numpy(), points[:total]


def mask_to_polygons(mask):
    """{mask id: [(2, N) array (rows y, x)]} for a (H,W) uint16 instance mask (NumPy or CUDA tensor)."""
    if isinstance(mask, torch.Tensor):
        dev_mask = mask
    else:
        m = np.ascontiguousarray(mask)
        if m.dtype != np.uint16:
            if m.min() < 0 or m.max() > 65535:
                raise ValueError("instance ids must fit in uint16")
            m = m.astype(np.uint16)
        if not torch.cuda.is_available():
            raise RuntimeError("microbeseg_b200.polygons needs a CUDA device (no CPU fallback)")
        dev_mask = torch.from_numpy(m.view(np.int16)).cuda()
    offsets, points = contours_device(dev_mask)
    pts = points.cpu().numpy().astype(np.int64)
    out = {}
    for l in range(len(offsets) - 1):
        if offsets[l + 1] > offsets[l]:
            out[l + 1] = [np.ascontiguousarray(pts[offsets[l]:offsets[l + 1]].T)]
    return out


def points_string(polygon_points):
    """infer.py:281-284: "x,y " per contour point"""
    return "".join("{},{} ".format(polygon_points[1, c], polygon_points[0, c]) for c in range(polygon_points.shape[1]))
