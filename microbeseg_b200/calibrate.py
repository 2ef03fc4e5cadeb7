"""Least-squares fit of the two 1x1 heads of a (random-init) DU net to synthetic distance maps.

No checkpoint or dataset is reachable offline and a random-init net predicts near-constant maps
(no pixel passes th_seed), which makes post-processing degenerate.  Fitting only the
2 x 65 head parameters (of 46 M) on synthetic frames whose instance masks are known gives cell-like
border / cell maps, so that tests and the benchmark exercise seeds, labelling and the watershed on
network outputs.  Everything runs on the CUDA path (features come from the fused last conv)."""
import numpy as np
import torch

from . import synthetic as sy


def synthetic_training_pair(H, W, seed, noise=0.0):
    """(uint16 frame, border target (H,W) f32, cell target (H,W) f32, instance mask)."""
    n_cells = max(int(H * W * 0.4 / 330.0), 1)
    mask = sy.synth_instance_mask(H, W, n_cells, seed + 1)
    frame = sy.synth_frame(H, W, seed)              # rendered from the same mask (same seed convention)
    border, cell = sy.synth_distance_maps(mask, seed + 2, noise=noise)
    return frame, border[..., 0], cell[..., 0], mask


@torch.no_grad()
def fit_heads(net, pairs, ridge=1e-3):
    """pairs: iterable of (frame uint16 (H,W) with sides multiple of 16, border target, cell target)."""
    dev = next(net.parameters()).device
    eng = net.engine()
    names = list(net.decoder_names)
    C = net._chans[0]
    A = {n: torch.zeros((C + 1, C + 1), dtype=torch.float64, device=dev) for n in names}
    rhs = {n: torch.zeros((C + 1,), dtype=torch.float64, device=dev) for n in names}
    for frame, border_t, cell_t in pairs:
        d = torch.from_numpy(np.ascontiguousarray(frame).view(np.int16)).to(dev)
        feats = {}
        eng.run(d[None], 0, 0, float(frame.min()), float(frame.max()), keep_features=feats)
        targets = {names[0]: border_t, names[-1]: cell_t}
        for n in names:
            F = feats[n][0][..., :C].reshape(-1, C).double()      # narrow nets run zero-padded to 64 channels
            F1 = torch.cat([F, torch.ones((F.shape[0], 1), dtype=torch.float64, device=dev)], 1)
            y = torch.from_numpy(np.ascontiguousarray(targets[n], dtype=np.float64)).to(dev).reshape(-1)
            A[n] += F1.T @ F1
            rhs[n] += F1.T @ y
    for n in names:
        reg = ridge * torch.eye(C + 1, dtype=torch.float64, device=dev) * A[n].diagonal().mean()
        reg[C, C] = 0
        sol = torch.linalg.solve(A[n] + reg, rhs[n])
        head = getattr(net, n + "Conv")[len(net._chans) - 1]
        head.weight.copy_(sol[:C].float().reshape(1, C, 1, 1))
        head.bias.copy_(sol[C:].float())
    return net


def train_briefly(net, steps=400, crop=256, n_crops=48, batch=8, seed0=9000, lr=8e-4, log=None):
    """Short CUDA training run on synthetic crops so that a random-init DU net predicts well-separated distance
    maps (tests / bench calibration only; no checkpoint or dataset is reachable offline).  Everything runs on this
    repo's own path: labels from ``labels.create_labels_device`` (the CUDA distance transforms), the step from
    ``training.TrainEngine`` (tcgen05 forward / backward) with the reference's Adam settings (train.py:380-385).
    Frames are normalised per frame to [-1, 1] exactly as the inference path does (infer.py:346).  Returns the
    list of losses; the net is left in eval() mode with its trained weights and BatchNorm running statistics."""
    from . import labels as lb
    from .training import TrainEngine, train_step
    dev = next(net.parameters()).device
    frames, masks = [], []
    for k in range(n_crops):
        frame, _, _, mask = synthetic_training_pair(crop, crop, seed0 + 10 * k)
        lo, hi = float(frame.min()), float(frame.max())
        frames.append(2 * (frame.astype(np.float32) - lo) / (hi - lo) - 1)
        masks.append(mask.astype(np.uint16))
    with torch.cuda.device(dev):
        md = torch.from_numpy(np.stack(masks).view(np.int16)).to(dev)
        cell, neigh, _ = lb.create_labels_device(md, int(np.stack(masks).max()))
        x = torch.from_numpy(np.stack(frames)[:, None]).to(dev)
        cell, neigh = cell[:, None].contiguous(), neigh[:, None].contiguous()
        net.train()
        eng = TrainEngine(net)
        opt = torch.optim.Adam(net.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=True)
        g = torch.Generator(device="cpu").manual_seed(seed0)
        losses = []
        for it in range(steps):
            idx = torch.randperm(n_crops, generator=g)[:batch].to(dev)
            loss = train_step(eng, opt, x[idx].contiguous(), neigh[idx].contiguous(), cell[idx].contiguous())
            if it % 50 == 0 or it == steps - 1:
                losses.append(float(loss))
                if log:
                    log(f"train_briefly step {it}: loss {losses[-1]:.5f}")
        net.eval()
    return losses
