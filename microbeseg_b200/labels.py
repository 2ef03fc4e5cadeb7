"""Host-side mirror of the reference's label generation for the distance method (CUDA path).

Same names / argument meaning as /root/reference/src/training/train_data_representations.py:
``get_label(mask, 'distance', max_mal)`` (:11-37) and ``distance_label(label, search_radius)``
(:261-361) return ``(cell_dist float32 (H,W), neighbor_dist float32 (H,W))``.  ``create_labels``
is the numerical part of ``CreateLabelsWorker.create_labels`` (src/training/train.py:63-96) for a
batch of crops: max_mal from the major axis lengths, then both maps, crops batched on the GPU.
No CPU fallback.
"""
import numpy as np
import torch

from . import _native as nat

_PINNED = {}
_MAX_BATCH_PIXELS = 1 << 28          # crops per C-ABI call are chunked to stay under 2^31 pixels / few GiB scratch


def _masks_to_device(masks, device):
    m = np.ascontiguousarray(masks)
    if m.ndim == 2:
        m = m[None]
    if m.dtype != np.uint16:
        if m.min() < 0 or m.max() > 65535:
            raise ValueError("instance ids must fit in uint16")
        m = m.astype(np.uint16)
    return torch.from_numpy(m.view(np.int16)).to(device), int(m.max())


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("microbeseg_b200.labels needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _run(masks_dev, max_id, search_radius, radius_hint):
    L = nat.lib()
    n, H, W = masks_dev.shape
    device = masks_dev.device
    cell = torch.empty((n, H, W), dtype=torch.float32, device=device)
    neigh = torch.empty((n, H, W), dtype=torch.float32, device=device)
    mal = torch.zeros(n, dtype=torch.int32, device=device)
    err = torch.zeros(n, dtype=torch.int32, device=device)
    ws = torch.empty(L.mbs_labels_workspace_bytes(n, H, W, max_id), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        nat.check(L.mbs_distance_labels(masks_dev.data_ptr(), n, H, W, max_id, int(search_radius), int(radius_hint),
                                        cell.data_ptr(), neigh.data_ptr(), mal.data_ptr(), err.data_ptr(),
                                        ws.data_ptr(), ws.numel(), nat.stream_ptr()), "distance_labels")
    e = err.cpu().numpy()
    if e.any():
        raise RuntimeError(f"distance_labels: crops {np.flatnonzero(e).tolist()[:8]} exceeded a device limit "
                           f"(bit0: instance window too large for shared memory, bit1: > 4096 gaps): {e[e != 0][:8]}")
    return cell, neigh, mal


def max_major_axis_lengths(masks):
    """int(ceil(max major_axis_length)) per crop (train.py:74-79)."""
    L = nat.lib()
    dev, max_id = _masks_to_device(masks, _device())
    n, H, W = dev.shape
    out = torch.zeros(n, dtype=torch.int32, device=dev.device)
    ws = torch.empty(L.mbs_labels_workspace_bytes(n, H, W, max_id), dtype=torch.uint8, device=dev.device)
    nat.check(L.mbs_labels_max_mal(dev.data_ptr(), n, H, W, max_id, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                   nat.stream_ptr()), "labels_max_mal")
    return out.cpu().numpy()


def distance_label(label, search_radius):
    """Cell and neighbor distance label creation (train_data_representations.py:261-361)."""
    dev, max_id = _masks_to_device(label, _device())
    cell, neigh, _ = _run(dev, max_id, int(search_radius), int(search_radius))
    return cell[0].cpu().numpy(), neigh[0].cpu().numpy()


def _simple_label(label, mode):
    L = nat.lib()
    dev, _ = _masks_to_device(label, _device())
    n, H, W = dev.shape
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev.device)
    with torch.cuda.device(dev.device):
        nat.check(L.mbs_boundary_border_labels(dev.data_ptr(), n, H, W, mode, out.data_ptr(), nat.stream_ptr()),
                  "boundary_border_labels")
    res = out.cpu().numpy()
    return res[0] if np.asarray(label).ndim == 2 else res


def boundary_label(label):
    """Boundary label image (train_data_representations.py:75-99): uint8, 1 nucleus, 2 boundary."""
    return _simple_label(label, 0)


def border_label(label):
    """Border label image (train_data_representations.py:102-126): uint8, 1 nucleus, 2 border between touching nuclei."""
    return _simple_label(label, 1)


def get_label(mask, label_type, max_mal):
    """Calculate training data representation / label (train_data_representations.py:11-37)."""
    if label_type == 'distance':
        return distance_label(mask, search_radius=int(np.ceil(0.75 * max_mal)))
    if label_type == 'boundary':
        return boundary_label(mask)
    if label_type == 'border':
        return border_label(mask)
    if label_type == 'cell_dist':
        # cell_distance_label(apply_clipping=False) (:220-258) is the cell half of distance_label: same windows, same
        # normalised EDTs, same `+=` (cells whose EDT is all zero add zeros there and are skipped here)
        return distance_label(mask, search_radius=int(np.ceil(0.75 * max_mal)))[0]
    if label_type in ('adapted_border', 'j4', 'cell_dist_clipped'):
        raise NotImplementedError(f"label type {label_type!r} is not built on the CUDA path "
                                  "('distance', 'cell_dist', 'boundary', 'border' are)")
    raise Exception('Label type not known')


def create_labels_device(masks_dev, max_id, radius_hint=-1):
    """Device-resident batch entry: int16-viewed uint16 masks [n,H,W] on the GPU -> (cell, neighbor, max_mal)."""
    return _run(masks_dev, max_id, -1, radius_hint)


def create_labels(masks):
    """Batch version of CreateLabelsWorker.create_labels (train.py:63-96) for [n,H,W] masks.
    Returns (cell_dist [n,H,W] f32, neighbor_dist [n,H,W] f32, max_mal [n] int)."""
    masks = np.asarray(masks)
    if masks.ndim == 2:
        masks = masks[None]
    n, H, W = masks.shape
    # <= 16 Mpx per batch: 2 x 64 MiB of pinned staging (measured end to end on 4000 crops of 320^2: 473 Mpx/s with
    # 16 Mpx batches, 284 with 64 Mpx, 161 with 256 Mpx -- allocation and first touch of large staging buffers dominate)
    per = max(1, min(n, (_MAX_BATCH_PIXELS // 16) // (H * W)))
    L = nat.lib()
    device = _device()
    cells = np.empty((n, H, W), np.float32)
    neighs = np.empty((n, H, W), np.float32)
    mals = np.empty(n, np.int32)
    # pinned staging for the read-back of one batch (two float32 maps per pixel dominate the host traffic)
    key = (per, H, W)
    if key not in _PINNED:
        _PINNED.clear()                                # keep one staging pair
        _PINNED[key] = (torch.empty((per, H, W), dtype=torch.float32, pin_memory=True),
                        torch.empty((per, H, W), dtype=torch.float32, pin_memory=True))
    pin_c, pin_n = _PINNED[key]
    for s in range(0, n, per):
        dev, max_id = _masks_to_device(masks[s:s + per], device)       # one upload per batch
        k = dev.shape[0]
        hint = 0
        if max_id > 0:       # radius hint from the batch's own max_mal, computed on the uploaded masks
            mal_dev = torch.zeros(k, dtype=torch.int32, device=device)
            ws = torch.empty(L.mbs_labels_workspace_bytes(k, H, W, max_id), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                nat.check(L.mbs_labels_max_mal(dev.data_ptr(), k, H, W, max_id, mal_dev.data_ptr(), ws.data_ptr(), ws.numel(),
                                               nat.stream_ptr()), "labels_max_mal")
            hint = int(np.ceil(0.75 * int(mal_dev.max().item())))
        c, nb, mal = _run(dev, max_id, -1, hint)
        pin_c[:k].copy_(c, non_blocking=True)
        pin_n[:k].copy_(nb, non_blocking=True)
        mals[s:s + k] = mal.cpu().numpy()                               # synchronises the stream
        cells[s:s + k] = pin_c[:k].numpy()
        neighs[s:s + k] = pin_n[:k].numpy()
    return cells, neighs, mals


def create_labels_sharded(masks, transport="auto"):
    """``create_labels`` under a torchrun launch: crop i -> rank i mod world, each rank runs its
    crops on its own GPU, rank 0 returns (cell_dist, neighbor_dist, max_mal) for all crops, the other ranks None.
    Mirrors the crop loop of CreateLabelsWorker.create_labels (train.py:63-96), whose iterations are independent."""
    from . import sharding
    masks = np.asarray(masks)
    if masks.ndim == 2:
        masks = masks[None]
    n, H, W = masks.shape

    def work(indices, outs):
        if len(indices) == 0:
            return
        c, nb, mal = create_labels(masks[indices])
        for j, t in enumerate(indices):
            outs[0][t], outs[1][t], outs[2][t] = c[j], nb[j], mal[j]

    res = sharding.run_sharded(work, n, [((H, W), np.float32), ((H, W), np.float32), ((), np.int32)], transport=transport)
    return None if res is None else tuple(res)
