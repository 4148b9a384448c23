"""``.pytorch`` model file -> B200 model container
(volume_segmantics/model/model_2d.py:10-57).

The file format is the contract: a ``torch.save`` dict with
``model_state_dict`` (smp key names), ``model_struc_dict`` {type: ModelType,
encoder_name, encoder_weights, in_channels, classes} and ``label_codes``
(early_stopping.py:55-62).  ``encoder_weights`` is ignored (no download)."""
import logging
from pathlib import Path
from typing import Tuple

import torch

from ..plan import B200SegmentationModel
from .enums import ModelType


def create_model_on_device(device_num: int, model_struc_dict: dict) -> torch.nn.Module:
    """model_2d.py:10-39.  The returned nn.Module holds the weights (on the host);
    the network itself runs in libvsb200 on GPU ``device_num``."""
    struct = dict(model_struc_dict)
    model_type = struct.pop("type")
    if not isinstance(model_type, ModelType):
        model_type = ModelType[str(model_type).upper()]
    struct.pop("encoder_weights", None)
    logging.info(f"Building the {model_type.name} model for the B200 engine on device {device_num}")
    model = B200SegmentationModel(
        model_type.name,
        struct.get("encoder_name", "resnet34"),
        int(struct.get("classes", 1)),
        int(struct.get("in_channels", 1)),
    )
    model.device_num = int(device_num)
    return model.eval()


def create_model_from_file(weights_fn: Path, gpu: bool = True, device_num: int = 0) -> Tuple[torch.nn.Module, int, dict]:
    """model_2d.py:42-57 -> (model, number of labels, label codes)."""
    weights_fn = Path(weights_fn).resolve()
    logging.info("Loading model dictionary from file.")
    # the dict pickles the ModelType enum, which torch >= 2.6's weights_only
    # default rejects; the file is the user's own checkpoint, as in the reference
    model_dict = torch.load(weights_fn, map_location="cpu", weights_only=False)
    model = create_model_on_device(device_num, model_dict["model_struc_dict"])
    logging.info("Loading in the saved weights.")
    model.load_state_dict(model_dict["model_state_dict"])
    return model, model_dict["model_struc_dict"]["classes"], model_dict["label_codes"]
