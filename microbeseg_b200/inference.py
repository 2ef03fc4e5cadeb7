"""Frame loop of the reference's inference entry points on the CUDA path.

Mirrors ``InferWorker.inference`` (src/inference/infer.py:328-376) and the per-frame loop of
``infer_script_local.py:118-161``: per-frame min/max -> top/left padding to a tested size ->
normalisation -> network -> crop of the pads -> distance post-processing -> uint16 mask.
Differences are purely mechanical: the frame is uploaded raw (uint8/uint16) from pinned memory,
normalisation + padding are fused into the first conv kernel, the distance maps never leave the
GPU, and only the uint16 mask comes back.  Frames of a 2D+t stack are independent (per-frame
min/max), so multi-GPU runs shard frames over ranks with no collective.
"""
import os

import numpy as np
import torch

from . import _native as nat
from . import postprocessing as pp
from .utils import model_input_pads


def shard_frames(num_frames, rank, world_size):
    """Frame indices handled by ``rank``: t = rank (mod world_size)."""
    return list(range(rank, num_frames, world_size))


class FrameSegmenter:
    """Reusable pinned/device staging for one frame size, double buffered.  H2D, network, post-processing and
    D2H run on four streams chained by events: the copy of frame k+1, the post-processing of frame k-1 (small
    latency-bound kernels that fit beside the persistent conv CTAs) and the read-back of frame k-2 overlap the
    network of frame k; frame min/max are reduced on the GPU (no host pass over the pixels)."""

    def __init__(self, net, ths, device=None):
        self.net = net
        self.th_cell, self.th_seed = float(ths[0]), float(ths[1])     # thresholds = [th_cell, th_seed]
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("microbeseg_b200.inference needs a CUDA device (no CPU fallback)")
        self._stage = {}
        self.h2d_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        # Post-processing runs on the network's stream.  The persistent conv CTAs fill every SM, so a second stream gives
        # no overlap (measured: 11.0 ms per frame either way), and the cooperative post-processing kernels -- which need the
        # whole GPU at once -- are starved behind the queued conv launches of the next frames until the host blocks on a
        # result, which leaves bubbles (e2e 344 vs 376 Mpx/s).  MBS_PP_STREAM=1 selects the separate stream (A/B knob).
        self.pp_stream = (torch.cuda.Stream(self.device) if os.environ.get("MBS_PP_STREAM") == "1"
                          else torch.cuda.current_stream(self.device))

    def _staging(self, shape, dtype, slot):
        key = (tuple(shape), np.dtype(dtype).str, slot)
        st = self._stage.get(key)
        if st is None:
            tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.int16,
                   np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
            st = dict(pin_in=torch.empty(shape, dtype=tdt).pin_memory(),
                      dev_in=torch.empty(shape, dtype=tdt, device=self.device),
                      dev_out=torch.empty(shape, dtype=torch.int16, device=self.device),
                      pin_out=torch.empty(shape, dtype=torch.int16).pin_memory(),
                      lohi=torch.empty(2, dtype=torch.float32, device=self.device),
                      scratch=torch.empty(2, dtype=torch.int32, device=self.device),
                      ev_h2d=torch.cuda.Event(), ev_net=torch.cuda.Event(), ev_done=torch.cuda.Event(),
                      ev_out=torch.cuda.Event())
            self._stage[key] = st
        return st

    @staticmethod
    def _canon(frame):
        frame = np.ascontiguousarray(frame)
        if frame.dtype not in (np.uint8, np.uint16, np.float32):
            frame = frame.astype(np.float32)
        return frame

    def submit(self, frame, slot=0, min_val=None, max_val=None, crop=None):
        """Enqueue H2D + network + post-processing + D2H for one (H,W) frame; returns a handle.
        ``crop=[py,px]``: the frame is already padded by the caller; crop the outputs by py/px."""
        from .unets import frame_minmax
        frame = self._canon(frame)
        H, W = frame.shape
        pads = model_input_pads(H, W) if crop is None else [0, 0]
        if len(pads) < 2:
            raise Exception('Image too big to pad. Use sliding windows')
        cy, cx = (0, 0) if crop is None else (int(crop[0]), int(crop[1]))
        st = self._staging((H, W), frame.dtype, slot)
        st["ev_h2d"].synchronize()                  # the previous upload from this pinned buffer is done
        st["pin_in"].copy_(torch.from_numpy(frame.view(np.int16) if frame.dtype == np.uint16 else frame))
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.h2d_stream):
                self.h2d_stream.wait_event(st["ev_net"])           # the network that read dev_in (2 frames ago) is done
                st["dev_in"].copy_(st["pin_in"], non_blocking=True)
                st["ev_h2d"].record(self.h2d_stream)
            main.wait_event(st["ev_h2d"])
            distance = len(getattr(self.net, "decoder_names", ())) == 2      # DU net = distance method, U net = boundary
            try:
                if min_val is None or max_val is None:
                    # frame min / max before padding (infer_script_local.py:124), reduced on the device
                    lohi = frame_minmax(st["dev_in"], out=st["lohi"], scratch=st["scratch"])
                    maps = self.net.forward_frame(st["dev_in"], pads, lohi_dev=lohi)
                else:
                    maps = self.net.forward_frame(st["dev_in"], pads, float(min_val), float(max_val))
            except RuntimeError:
                # same contract as infer.py:352-356: a RuntimeError during net() yields an empty mask
                maps = None
                print('RuntimeError during inference (maybe not enough ram/vram?)')
            st["ev_net"].record(main)
            with torch.cuda.stream(self.pp_stream):
                self.pp_stream.wait_event(st["ev_net"])
                self.pp_stream.wait_event(st["ev_out"])            # read-back of this slot's previous mask is done
                if maps is None:
                    st["dev_out"].zero_()
                    st["view"] = (cy, cx) if (cy, cx) != (0, 0) else None
                else:
                    if distance:
                        border, cell = maps
                        border.record_stream(self.pp_stream)
                        cell.record_stream(self.pp_stream)
                        b = border[0, 0, pads[0] + cy:, pads[1] + cx:]      # crop the pads (infer.py:358-359)
                        c = cell[0, 0, pads[0] + cy:, pads[1] + cx:]
                        run = lambda dst: pp.distance_postprocessing_device(b, c, self.th_seed, self.th_cell, out=dst)
                    else:
                        # boundary method: softmax over the 3 classes, channel-last crop (infer.py:371-374)
                        maps.record_stream(self.pp_stream)
                        hp, wp = maps.shape[-2], maps.shape[-1]
                        y0, x0 = pads[0] + cy, pads[1] + cx
                        prob = torch.empty((hp - y0, wp - x0, 3), dtype=torch.float32, device=maps.device)
                        nat.check(nat.lib().mbs_softmax3_hwc(maps.data_ptr(), hp * wp, wp, y0, x0, hp - y0, wp - x0, prob.data_ptr(),
                                                             nat.stream_ptr(self.pp_stream)), "softmax3_hwc")
                        run = lambda dst: pp.boundary_postprocessing_device(prob, out=dst)
                    if (cy, cx) == (0, 0):
                        run(st["dev_out"])
                        st["view"] = None
                    else:
                        small = run(None)
                        st["dev_out"].zero_()
                        st["dev_out"][cy:, cx:] = small
                        st["view"] = (cy, cx)
                st["ev_done"].record(self.pp_stream)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(st["ev_done"])
                st["pin_out"].copy_(st["dev_out"], non_blocking=True)
                st["ev_out"].record(self.d2h_stream)
        return st

    @staticmethod
    def result(handle, out=None):
        """Mask of a submitted frame.  ``out``: caller-provided (H,W) uint16 array the mask is written into straight
        from the pinned read-back buffer (one host copy instead of two); default: a fresh array."""
        handle["ev_out"].synchronize()
        res = handle["pin_out"].numpy().view(np.uint16)
        v = handle.get("view")
        if isinstance(v, tuple):
            res = res[v[0]:, v[1]:]
        if out is None:
            return res.copy()
        np.copyto(out, res)
        return out

    def segment(self, frame, min_val=None, max_val=None, crop=None):
        return self.result(self.submit(frame, 0, min_val, max_val, crop))


TILE_HALO = 128      # >= the network's 116-px receptive-field reach, multiple of 16


def predict_maps_tiled(net, frame_dev, lo, hi, tile=1024, halo=TILE_HALO):
    """Overlap-tiled network inference that is bit-identical to whole-frame inference.

    The reference only has whole-frame inference (its ``sliding_window`` flag is a stub,
    src/inference/infer.py:60,76) and refuses frames above 8192 px (src/utils/utils.py:154-155).
    Tiles have 16-aligned origins and a ``halo``-pixel apron: every kept output pixel sees exactly the
    inputs it sees in the whole frame (true zero padding only at real image borders), so stitching the
    tile cores reproduces the whole-frame maps bit for bit.  ``frame_dev``: raw (H,W) CUDA frame whose
    sides are multiples of 16 (already padded); returns (border, cell) float32 (H,W)."""
    H, W = frame_dev.shape
    if H % 16 or W % 16 or tile % 16 or halo % 16:
        raise RuntimeError("tiled inference needs frame sides, tile and halo to be multiples of 16")
    border = torch.empty((H, W), dtype=torch.float32, device=frame_dev.device)
    cell = torch.empty((H, W), dtype=torch.float32, device=frame_dev.device)
    for y0 in range(0, H, tile):
        for x0 in range(0, W, tile):
            y1, x1 = min(y0 + tile, H), min(x0 + tile, W)
            wy0, wy1 = max(y0 - halo, 0), min(y1 + halo, H)
            wx0, wx1 = max(x0 - halo, 0), min(x1 + halo, W)
            win = frame_dev[wy0:wy1, wx0:wx1].contiguous()
            b, c = net.forward_frame(win, [0, 0], float(lo), float(hi))
            border[y0:y1, x0:x1] = b[0, 0, y0 - wy0:y1 - wy0, x0 - wx0:x1 - wx0]
            cell[y0:y1, x0:x1] = c[0, 0, y0 - wy0:y1 - wy0, x0 - wx0:x1 - wx0]
    return border, cell


def segment_frame_tiled(net, frame, ths=(0.10, 0.45), tile=1024, device=None):
    """Whole path for one large frame with overlap-tiled inference: min/max -> top/left padding with the
    frame minimum -> tiled network -> crop -> distance post-processing."""
    device = torch.device(device) if device is not None else next(net.parameters()).device
    frame = FrameSegmenter._canon(frame)
    H, W = frame.shape
    lo, hi = frame.min(), frame.max()
    pads = model_input_pads(H, W)
    # same top/left padding as the reference while the frame fits its table (<= 8192); beyond it the
    # reference raises ("Use sliding windows"), here the frame is padded to the next multiple of 16
    py, px = (pads[0], pads[1]) if len(pads) == 2 else ((-H) % 16, (-W) % 16)
    padded = np.pad(frame, ((py, 0), (px, 0)), mode="constant", constant_values=lo)
    dev = torch.from_numpy(padded.view(np.int16) if padded.dtype == np.uint16 else padded).to(device)
    with torch.cuda.device(device):
        border, cell = predict_maps_tiled(net, dev, lo, hi, tile)
        out = pp.distance_postprocessing_device(border[py:, px:], cell[py:, px:], float(ths[1]), float(ths[0]))
    return out.cpu().numpy().view(np.uint16)


def segment_stack(net, stack, ths=(0.10, 0.45), device=None, frames=None, out=None):
    """[T,H,W] stack -> [T,H,W] uint16 masks (infer_script_local.py:115-161).  ``frames`` selects the
    frame indices to process (frame sharding); other rows of ``out`` are left untouched."""
    stack = np.asarray(stack)
    if stack.ndim == 2:
        stack = stack[None]
    T = stack.shape[0]
    if out is None:
        from . import staging
        out = staging.host_empty(stack.shape, np.uint16, zero=True)      # large stacks: anonymous mapping with 2 MiB pages
    seg = FrameSegmenter(net, ths, device)
    todo = list(range(T)) if frames is None else list(frames)
    pending = None
    for k, t in enumerate(todo):
        h = seg.submit(stack[t], slot=k & 1)
        if pending is not None:
            seg.result(pending[1], out=out[pending[0]])
        pending = (t, h)
    if pending is not None:
        seg.result(pending[1], out=out[pending[0]])
    return out


def segment_stack_sharded(net, stack, ths=(0.10, 0.45), device=None, transport="auto"):
    """``segment_stack`` under a torchrun launch: frame t -> rank t mod world, every rank runs its frames on its own
    GPU (weights replicated, no collective on the data path), rank 0 receives all masks and returns the full
    [T,H,W] uint16 array (the other ranks return None).  Without a process group it is plain ``segment_stack``."""
    from . import sharding
    stack = np.asarray(stack)
    if stack.ndim == 2:
        stack = stack[None]
    T, H, W = stack.shape

    def work(indices, outs):
        segment_stack(net, stack, ths=ths, device=device, frames=indices, out=outs[0])

    res = sharding.run_sharded(work, T, [((H, W), np.uint16)], transport=transport)
    return None if res is None else res[0]


class InferWorker:
    """The numerical part of the reference's ``InferWorker`` (src/inference/infer.py:30-94, 328-376):
    same constructor argument order; OMERO transport / Qt signals are out of scope."""

    def __init__(self, img_id_list, inference_path, omero_username, omero_password, omero_host, omero_port, group_id,
                 model, device, ths, channel=0, upload=True, overwrite=True, sliding_window=False,
                 print_output=False):
        import json
        from pathlib import Path
        from .unets import build_unet, get_weights
        self.img_id_list, self.inference_path, self.device, self.ths = img_id_list, inference_path, device, ths
        self.channel, self.upload, self.overwrite, self.print_output = channel, upload, overwrite, print_output
        self.model = Path(model)
        with open(self.model.parent / f"{self.model.stem}.json") as f:
            self.model_settings = json.load(f)
        arch = self.model_settings['architecture']
        self.net = build_unet(unet_type=arch[0], act_fun=arch[2], pool_method=arch[1], normalization=arch[3],
                              device=device, num_gpus=1, ch_in=1,
                              ch_out=1 if self.model_settings['label_type'] == 'distance' else 3, filters=arch[4])
        self.net = get_weights(net=self.net, weights=str(self.model.parent / f"{self.model.stem}.pth"),
                               num_gpus=1, device=device)
        self.net.eval()
        self._seg = None

    def inference(self, img, min_val, max_val, pads):
        """``img`` is the already padded frame (infer.py:256-259); returns the uint16 mask without pads."""
        torch.set_grad_enabled(False)
        if self._seg is None:
            self._seg = FrameSegmenter(self.net, self.ths, self.device)
        # the caller's padded frame is used as is (whatever its pad value); only the crop differs
        return self._seg.segment(np.asarray(img), min_val, max_val, crop=pads)
