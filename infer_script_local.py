"""microbeSEG inference script for local .tif stacks -- same command line as the reference's
infer_script_local.py (:17-25), running on the sm_100a CUDA path (no CPU mode)."""
import argparse
import json
from pathlib import Path

import numpy as np
import torch

from microbeseg_b200 import tiffio as tiff
from microbeseg_b200 import sharding
from microbeseg_b200.inference import segment_stack_sharded
from src.utils.unets import build_unet, get_weights


def main(argv=None):
    parser = argparse.ArgumentParser(description='microbeSEG inference script')
    parser.add_argument('--img_dir', '-i', required=True, type=str, help='Directory with image files to process (.tif, .tiff)')
    parser.add_argument('--model', '-m', required=True, type=str, help='Path to model')
    parser.add_argument('--thresholds', '-t', default=[0.10, 0.45], type=float, nargs='+', help='Thresholds for distance models')
    parser.add_argument('--result_path', '-r', default=None, type=str, help='Result path')
    parser.add_argument('--channel', '-c', default=0, type=int, help='Channel to process')
    parser.add_argument('--device', '-d', default='cuda:0', help='Device to use')
    parser.add_argument('--overwrite', '-o', default=False, action='store_true', help='Overwrite existing results')
    args = parser.parse_args(argv)

    imgs_path = Path(args.img_dir)
    result_path = (Path(__file__).parent / 'results') if args.result_path is None else Path(args.result_path)
    result_path.mkdir(exist_ok=True)

    inference_model = Path(args.model)
    for ext in ('.pth', '.json'):
        if not (inference_model.parent / f"{inference_model.stem}{ext}").is_file():
            raise Exception(f'{inference_model.parent / f"{inference_model.stem}{ext}"} not found!')
    with open(inference_model.parent / f"{inference_model.stem}.json") as f:
        model_settings = json.load(f)
    if len(args.thresholds) != 2:
        raise Exception(f"{len(args.thresholds)} threshold given, needed are 2")
    if 'cuda' not in args.device:
        raise ValueError('this build runs on CUDA devices only (no CPU path); use the reference for device "cpu"')
    if not torch.cuda.is_available():
        raise ValueError('No cuda capable gpu device detected')
    # under torchrun (one process per GPU) the frames of every stack are sharded over the ranks, frame t -> rank
    # t mod world; rank 0 gathers the masks and is the only writer; plain launch: one GPU as given
    rank, world, local_rank = sharding.init_from_env()
    device = torch.device('cuda', local_rank) if world > 1 else torch.device(args.device)
    say = print if rank == 0 else (lambda *a, **k: None)

    file_ids = sorted(imgs_path.glob('*.tif*'))
    if len(file_ids) == 0:
        say('No files found')
        return

    arch = model_settings['architecture']
    net = build_unet(unet_type=arch[0], act_fun=arch[2], pool_method=arch[1], normalization=arch[3], device=device,
                     num_gpus=1, ch_in=1, ch_out=1 if model_settings['label_type'] == 'distance' else 3, filters=arch[4])
    net = get_weights(net=net, weights=str(inference_model.parent / f"{inference_model.stem}.pth"), num_gpus=1, device=device)
    net.eval()
    torch.set_grad_enabled(False)
    say('--- Start inference ---')
    for img_id in file_ids:
        img = tiff.imread(str(img_id))
        fname = result_path / img_id.stem
        # image needs to be in shape [time dimension, height, width] (infer_script_local.py:86-101)
        if img.ndim == 2:
            img = img[None, ...]
        elif img.ndim == 3:
            if img.shape[-1] == 3:
                img = img[..., args.channel][None, ...]
            elif img.shape[0] == 3:
                img = img[args.channel, ...][None, ...]
        elif img.ndim == 4:
            img = img[:, args.channel, ...]
        elif img.ndim == 5:
            say(f'Skip {fname.name} (not supported image shape)')
            continue
        else:
            raise Exception('Adapt script for your data format!')
        out_file = result_path / f"mask_{fname.stem}_channel{args.channel}.tif"
        if out_file.is_file() and not args.overwrite:
            say(f'Skip {fname.name} (already processed and overwriting not enabled)')
            continue
        say(f'Process {fname.name} (channel: {args.channel})')
        results_array = segment_stack_sharded(net, img, ths=args.thresholds, device=device)
        if rank == 0:
            tiff.imwrite(str(out_file), np.squeeze(results_array))
    say('--- Finished ---')
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
