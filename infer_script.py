"""OMERO inference entry point of the reference (infer_script.py:15-30): same command line.

The OMERO transport (omero-py BlitzGateway, ROI upload) is outside the hot-path scope;
the numerical part it drives is ``microbeseg_b200.inference.InferWorker.inference`` and the local-file CLI
``infer_script_local.py``.  This stub keeps the argument surface and explains how to proceed."""
import argparse


def main(argv=None):
    parser = argparse.ArgumentParser(description='microbeSEG inference script (OMERO)')
    parser.add_argument('--ids', '-i', required=True, type=int, nargs='+', help='OMERO project/dataset/image ids')
    parser.add_argument('--id_type', '-it', required=True, type=str, help='"project", "dataset", "image"')
    parser.add_argument('--model', '-m', required=True, type=str, help='Path to model')
    parser.add_argument('--username', '-u', required=True, type=str, help='OMERO username')
    parser.add_argument('--password', '-p', required=True, type=str, help='OMERO password')
    parser.add_argument('--thresholds', '-t', default=[0.10, 0.45], type=float, nargs='+', help='Thresholds')
    parser.add_argument('--channel', '-c', default=0, type=int, help='Channel to process')
    parser.add_argument('--device', '-d', default='cuda:0', help='Device to use')
    parser.add_argument('--group_id', '-g', default=None, type=int, help='OMERO group id')
    parser.add_argument('--upload', default=False, action='store_true', help='Upload results to OMERO')
    parser.add_argument('--overwrite', '-o', default=False, action='store_true', help='Overwrite existing results')
    parser.parse_args(argv)
    raise SystemExit("OMERO transport is not part of this build: export the images as .tif and run "
                     "infer_script_local.py, or drive microbeseg_b200.inference.InferWorker.inference() from your "
                     "own OMERO client.")


if __name__ == "__main__":
    main()
