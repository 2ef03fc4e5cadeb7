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


def _check_err(err):
    e = err.cpu().numpy()
    if e.any():
        raise RuntimeError(f"distance_labels: crops {np.flatnonzero(e).tolist()[:8]} exceeded a device limit "
                           f"(bit0: an instance wider than 58 px whose bounding box (+6) does not fit the closing kernel's shared memory, bit1: > 4096 gaps): {e[e != 0][:8]}")


def _run(masks_dev, max_id, search_radius, radius_hint, err=None, cell_clip=0.0):
    """``err``: int32 [n] device tensor that collects the per-crop limit flags (checked later by the caller, no
    synchronisation here); None = check immediately."""
    L = nat.lib()
    n, H, W = masks_dev.shape
    device = masks_dev.device
    cell = torch.empty((n, H, W), dtype=torch.float32, device=device)
    neigh = torch.empty((n, H, W), dtype=torch.float32, device=device)
    mal = torch.zeros(n, dtype=torch.int32, device=device)
    deferred = err is not None
    if err is None:
        err = torch.zeros(n, dtype=torch.int32, device=device)
    ws = torch.empty(L.mbs_labels_workspace_bytes(n, H, W, max_id), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        nat.check(L.mbs_distance_labels_ex(masks_dev.data_ptr(), n, H, W, max_id, int(search_radius), int(radius_hint),
                                           float(cell_clip), cell.data_ptr(), neigh.data_ptr(), mal.data_ptr(), err.data_ptr(),
                                           ws.data_ptr(), ws.numel(), nat.stream_ptr()), "distance_labels")
    if not deferred:
        _check_err(err)
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


def j4_label(label, k_neighbors=2, se_radius=4):
    """Pena / J4 label image (train_data_representations.py:158-190): uint8, 0 background, 1 cell, 2 touching, 3 gap."""
    L = nat.lib()
    dev, _ = _masks_to_device(label, _device())
    n, H, W = dev.shape
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev.device)
    tmp = torch.empty((n, H, W), dtype=torch.uint8, device=dev.device)
    with torch.cuda.device(dev.device):
        nat.check(L.mbs_j4_labels(dev.data_ptr(), n, H, W, int(k_neighbors), int(se_radius), out.data_ptr(), tmp.data_ptr(),
                                  nat.stream_ptr()), "j4_labels")
    res = out.cpu().numpy()
    return res[0] if np.asarray(label).ndim == 2 else res


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
    if label_type == 'cell_dist_clipped':
        # cell_distance_label(apply_clipping=True) (:246-256): the raw per-instance EDTs, clipped to clip_val = 5 and scaled
        dev, max_id = _masks_to_device(mask, _device())
        r = int(np.ceil(0.75 * max_mal))
        return _run(dev, max_id, r, r, cell_clip=5.0)[0][0].cpu().numpy()
    if label_type == 'j4':
        return j4_label(mask)
    if label_type == 'adapted_border':
        raise NotImplementedError("label type 'adapted_border' (cv2.Canny) is not built on the CUDA path "
                                  "('distance', 'cell_dist', 'cell_dist_clipped', 'boundary', 'border', 'j4' are)")
    raise Exception('Label type not known')


def create_labels_device(masks_dev, max_id, radius_hint=-1):
    """Device-resident batch entry: int16-viewed uint16 masks [n,H,W] on the GPU -> (cell, neighbor, max_mal)."""
    return _run(masks_dev, max_id, -1, radius_hint)


def create_labels(masks, out=None):
    """Batch version of CreateLabelsWorker.create_labels (train.py:63-96) for [n,H,W] masks.
    Returns (cell_dist [n,H,W] f32, neighbor_dist [n,H,W] f32, max_mal [n] int).

    Host path: masks go up and maps come back through pinned, chunked, multi-threaded staging (``staging``), batch
    k+1 is uploaded while batch k computes, and nothing synchronises per batch except the final read-back of a batch's
    maps.  ``out=(cell, neigh)``: caller-provided float32 [n,H,W] destinations -- NumPy arrays, or PINNED torch tensors,
    which are written by the copy engine directly (no host-side copy at all)."""
    from . import staging
    masks = np.asarray(masks)
    if masks.ndim == 2:
        masks = masks[None]
    if masks.dtype != np.uint16:
        if masks.min() < 0 or masks.max() > 65535:
            raise ValueError("instance ids must fit in uint16")
        masks = masks.astype(np.uint16)
    masks = np.ascontiguousarray(masks)
    n, H, W = masks.shape
    # <= 16 Mpx per batch (allocation and first touch of larger device / staging buffers dominate)
    per = max(1, min(n, (_MAX_BATCH_PIXELS // 16) // (H * W)))
    L = nat.lib()
    device = _device()
    pinned_out = out is not None and all(isinstance(o, torch.Tensor) and o.is_pinned() for o in out)
    if out is None:
        cells, neighs = staging.host_empty((n, H, W), np.float32), staging.host_empty((n, H, W), np.float32)
    else:
        cells, neighs = out
        for o in (cells, neighs):
            if tuple(o.shape) != (n, H, W) or (o.dtype not in (np.float32, torch.float32)):
                raise ValueError("create_labels: out arrays must be float32 [n,H,W]")
    mals_dev = torch.zeros(n, dtype=torch.int32, device=device)
    err_dev = torch.zeros(n, dtype=torch.int32, device=device)
    max_id_all = int(masks.max()) if n else 0
    starts = list(range(0, n, per))
    with torch.cuda.device(device):
        main = torch.cuda.current_stream(device)
        up, dn = torch.cuda.Stream(device), torch.cuda.Stream(device)
        bufs = [torch.empty((per, H, W), dtype=torch.int16, device=device) for _ in range(2)]
        hint_pin = torch.zeros(2, dtype=torch.int32).pin_memory()
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def upload(bi):
            """stream `up`: masks of batch bi -> device, and the batch's max major axis (sizes the EDT launch) -> pinned"""
            s0 = starts[bi]
            k = min(per, n - s0)
            with torch.cuda.stream(up):
                if bi >= 2:
                    up.wait_event(ev_free[bi & 1])               # the batch that used this buffer has been computed
                staging.upload(masks[s0:s0 + k].view(np.int16), bufs[bi & 1][:k])
                if max_id_all > 0:
                    mal_dev = torch.zeros(k, dtype=torch.int32, device=device)
                    ws = torch.empty(L.mbs_labels_workspace_bytes(k, H, W, max_id_all), dtype=torch.uint8, device=device)
                    nat.check(L.mbs_labels_max_mal(bufs[bi & 1][:k].data_ptr(), k, H, W, max_id_all, mal_dev.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), nat.stream_ptr()), "labels_max_mal")
                    hint_pin[bi & 1:(bi & 1) + 1].copy_(mal_dev.max().reshape(1), non_blocking=True)
                ev_up[bi & 1].record(up)

        def read_back(item):
            """stream `dn`: maps of a finished batch -> host (blocks the host, overlaps the next batch's kernels)"""
            c, nb, s0, k, ev = item
            with torch.cuda.stream(dn):
                dn.wait_event(ev)
                c.record_stream(dn)
                nb.record_stream(dn)
                if pinned_out:
                    cells[s0:s0 + k].copy_(c, non_blocking=True)
                    neighs[s0:s0 + k].copy_(nb, non_blocking=True)
                else:
                    staging.download(c, cells[s0:s0 + k] if isinstance(cells, np.ndarray) else cells[s0:s0 + k].numpy())
                    staging.download(nb, neighs[s0:s0 + k] if isinstance(neighs, np.ndarray) else neighs[s0:s0 + k].numpy())

        if starts:
            upload(0)
        pending = None
        for bi, s0 in enumerate(starts):
            k = min(per, n - s0)
            if bi + 1 < len(starts):
                upload(bi + 1)
            ev_up[bi & 1].synchronize()                          # masks + hint of this batch are there (uploaded a batch ago)
            main.wait_event(ev_up[bi & 1])
            hint = int(np.ceil(0.75 * int(hint_pin[bi & 1]))) if max_id_all > 0 else 0
            c, nb, mal = _run(bufs[bi & 1][:k], max_id_all, -1, hint, err=err_dev[s0:s0 + k])
            ev_free[bi & 1].record(main)
            mals_dev[s0:s0 + k] = mal
            ev_done = torch.cuda.Event()
            ev_done.record(main)
            if pending is not None:
                read_back(pending)
            pending = (c, nb, s0, k, ev_done)
        if pending is not None:
            read_back(pending)
        dn.synchronize()
        mals = mals_dev.cpu().numpy()
        _check_err(err_dev)
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
