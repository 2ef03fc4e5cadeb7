"""Host-side mirror of the reference's post-processing operators, running on libmbseg (CUDA).

Same names, argument order and return types as /root/reference/src/inference/postprocessing.py:
``distance_postprocessing(border_prediction, cell_prediction, th_seed, th_cell)`` (:7-59).
Inputs may be NumPy arrays of shape (H,W,1) / (H,W) (what the reference callers pass after
``.cpu().numpy()``, src/inference/infer.py:358-363) or CUDA float32 tensors (new: skips the D2H).
The result is a fresh ``np.uint16`` array of shape (H,W); inputs are never modified.
There is no CPU fallback: without a CUDA device or libmbseg.so a RuntimeError is raised.
"""
import ctypes

import numpy as np
import torch

from . import _native as nat
from . import staging

_WS = {}       # (device index, stream handle) -> workspace tensor
last_info = {}


def _workspace(device, nbytes):
    """Scratch for one call, cached per (device, stream): kernels of two streams never share a buffer, and a buffer
    is only ever reused by later work on the stream it was allocated on (stream order = reuse order), so neither the
    caching allocator nor a concurrent caller can hand it to someone else while kernels still use it."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if len(_WS) > 16:                     # streams come and go (FrameSegmenter instances): keep the cache small
            _WS.clear()
        with torch.cuda.device(device):
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _info_dict(info):
    return dict(n_seed_components=int(info[0]), n_markers=int(info[1]), sweeps=int(info[2]), sequential=int(info[3]),
                ambiguous=int(info[4]), overflow=int(info[5]), whole_image_sequential=int(info[6]),
                components_reflooded=int(info[7]))


def _as_device_map(a, device):
    """-> contiguous float32 CUDA tensor (H,W) without modifying the caller's array."""
    if isinstance(a, torch.Tensor):
        t = a
        if t.dim() == 3 and t.shape[-1] == 1:
            t = t[..., 0]
        if t.dim() != 2:
            raise ValueError(f"expected (H,W) or (H,W,1), got {tuple(a.shape)}")
        return t.to(device=device, dtype=torch.float32)
    arr = np.asarray(a)
    if arr.ndim == 3 and arr.shape[-1] == 1:
        arr = arr[..., 0]
    if arr.ndim != 2:
        raise ValueError(f"expected (H,W) or (H,W,1), got {arr.shape}")
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    with torch.cuda.device(device):
        t = torch.empty(arr.shape, dtype=torch.float32, device=device)
        staging.upload(arr, t)          # pinned, chunked, multi-threaded (a pageable copy runs at ~5 GB/s)
    return t


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("microbeseg_b200.postprocessing needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def distance_postprocessing_device(border, cell, th_seed, th_cell, out=None, want_info=False):
    """CUDA tensors in (float32, (H,W), row stride = ld elements), uint16-as-int16 tensor out (H,W)."""
    L = nat.lib()
    H, W = cell.shape
    if border.shape != cell.shape:
        raise ValueError("border and cell prediction shapes differ")
    if cell.stride(1) != 1 or border.stride(1) != 1 or cell.stride(0) != border.stride(0):
        cell, border = cell.contiguous(), border.contiguous()
    ld = cell.stride(0) if H > 1 else W
    device = cell.device
    if out is None:
        out = torch.empty((H, W), dtype=torch.int16, device=device)  # uint16 payload
    nbytes = L.mbs_postproc_workspace_bytes(H, W)
    info = (ctypes.c_int64 * 8)() if want_info else None
    with torch.cuda.device(device):
        ws = _workspace(device, nbytes)
        rc = L.mbs_distance_postprocessing(border.data_ptr(), cell.data_ptr(), H, W, ld, float(th_seed),
                                           float(th_cell), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                           ctypes.cast(info, ctypes.c_void_p) if want_info else None,
                                           nat.stream_ptr())
    nat.check(rc, "distance_postprocessing")
    if want_info:
        last_info.clear()
        last_info.update(_info_dict(info))
    return out


def distance_postprocessing(border_prediction, cell_prediction, th_seed, th_cell):
    """Post-processing for distance label (cell + neighbor) prediction -> uint16 instance mask (H,W).

    Mirrors postprocessing.py:7-59 (note the positional order: border, cell, th_seed, th_cell)."""
    device = _device_of(border_prediction, cell_prediction)
    border = _as_device_map(border_prediction, device)
    cell = _as_device_map(cell_prediction, device)
    out = distance_postprocessing_device(border, cell, th_seed, th_cell)
    res = staging.host_empty(tuple(out.shape), np.uint16)
    with torch.cuda.device(device):
        staging.download(out, res)
    # np.squeeze as in postprocessing.py:59 (a 1xW frame comes back 1-D, exactly like the reference)
    return np.squeeze(res)


def boundary_postprocessing_device(pred, out=None, want_info=False):
    """(H,W,3) float32 CUDA probabilities -> uint16-as-int16 CUDA mask (H,W)."""
    if pred.dim() != 3 or pred.shape[-1] != 3:
        raise ValueError(f"expected (H,W,3) class probabilities, got {tuple(pred.shape)}")
    pred = pred.contiguous()
    L = nat.lib()
    H, W = pred.shape[:2]
    device = pred.device
    if out is None:
        out = torch.empty((H, W), dtype=torch.int16, device=device)
    info = (ctypes.c_int64 * 8)() if want_info else None
    with torch.cuda.device(device):
        ws = _workspace(device, L.mbs_postproc_workspace_bytes(H, W))
        rc = L.mbs_boundary_postprocessing(pred.data_ptr(), H, W, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                           ctypes.cast(info, ctypes.c_void_p) if want_info else None, nat.stream_ptr())
    nat.check(rc, "boundary_postprocessing")
    if want_info:
        last_info.clear()
        last_info.update(_info_dict(info))
    return out


def boundary_postprocessing(prediction):
    """Post-processing for boundary label prediction -> uint16 instance mask (postprocessing.py:62-90).

    ``prediction``: (H,W,3) softmax probabilities (NumPy or CUDA tensor), channel-last as in the reference."""
    device = _device_of(prediction)
    if isinstance(prediction, torch.Tensor):
        pred = prediction.to(device=device, dtype=torch.float32)
    else:
        pred = torch.from_numpy(np.ascontiguousarray(prediction, dtype=np.float32)).to(device)
    out = boundary_postprocessing_device(pred)
    return np.squeeze(out.cpu().numpy().view(np.uint16))
