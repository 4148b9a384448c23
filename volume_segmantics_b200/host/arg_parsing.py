"""Command-line parsing for model-predict-2d
(volume_segmantics/utilities/arg_parsing.py:9-34, 83-120)."""
import argparse
from pathlib import Path

from . import constants as cfg


def CheckExt(choices):
    """argparse action factory: the file must exist and carry one of `choices` suffixes."""

    class Act(argparse.Action):
        def __call__(self, parser, namespace, fname, option_string=None):
            suffix = Path(fname).suffix
            if suffix not in choices:
                parser.error(f"Wrong filetype: file doesn't end with {choices}")
            if not Path(fname).is_file():
                parser.error(f"The file {str(fname)} does not appear to exist.")
            setattr(namespace, self.dest, fname)

    return Act


def get_2d_prediction_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(
        usage="%(prog)s path/to/model/file.zip path/to/data/file [path/to/data_directory]",
        description="Predict segmentation of a 3d data volume using the 2d model provided.",
    )
    parser.add_argument("-v", "--version", action="version", version=f"{parser.prog} version 1.0.0")
    parser.add_argument(cfg.MODEL_PTH_ARG, metavar="Model file path", type=str,
                        action=CheckExt(cfg.MODEL_DATA_EXT),
                        help="the path to a zip file containing the model weights.")
    parser.add_argument(cfg.PREDICT_DATA_ARG, metavar="Path to prediction data volume", type=str,
                        action=CheckExt(cfg.PREDICT_DATA_EXT),
                        help="the path to an HDF5 file containing the imaging data to segment")
    parser.add_argument("--" + cfg.DATA_DIR_ARG, metavar="Path to settings and output directory (optional)",
                        type=str, nargs="?", default=Path.cwd(),
                        help='path to a directory containing the "volseg-settings", data will be also be output to this location')
    return parser
