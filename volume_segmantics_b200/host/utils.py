"""Host utilities on either side of the prediction path; behaviour follows
volume_segmantics/utilities/base_data_utils.py (line numbers cited per function).
Volume I/O needs h5py / imageio exactly like the reference; they are imported
lazily so the GPU path itself has no such dependency."""
from __future__ import annotations

import logging
import sys
from enum import Enum
from pathlib import Path
from types import SimpleNamespace

import numpy as np

from . import constants as cfg
from .enums import Axis, ModelType, Quality


def create_enum_from_setting(setting_str, enum):
    """base_data_utils.py:53-64 -- invalid names log an error and exit(1)."""
    if isinstance(setting_str, Enum):
        return setting_str
    try:
        return enum[setting_str.upper()]
    except KeyError:
        options = [k.name for k in enum]
        logging.error(f"{enum.__name__}: {setting_str} is not valid. Options are {options}.")
        sys.exit(1)


def get_prediction_quality(settings: SimpleNamespace) -> Enum:
    return create_enum_from_setting(settings.quality, Quality)  # :67-69


def get_model_type(settings: SimpleNamespace) -> Enum:
    return create_enum_from_setting(settings.model["type"], ModelType)  # :72-74


def get_training_axis(settings: SimpleNamespace) -> Enum:
    return create_enum_from_setting(getattr(settings, "training_axes", "All"), Axis)  # :77-83


def get_prediction_axis(settings: SimpleNamespace) -> Enum:
    return create_enum_from_setting(getattr(settings, "prediction_axis", "Z"), Axis)  # :86-92


def setup_path_if_exists(input_param):
    """:95-101"""
    if isinstance(input_param, str):
        return Path(input_param)
    if isinstance(input_param, Path):
        return input_param
    return None


def get_batch_size(settings: SimpleNamespace, prediction: bool = False) -> int:
    """:104-122.  Kept for API compatibility; the B200 engine sizes its own
    slice batches (vsb_set_batch), this value is informational only."""
    import torch

    dev = int(settings.cuda_device)
    total = torch.cuda.get_device_properties(dev).total_memory
    free_gb = (total - torch.cuda.memory_allocated(dev)) / 1024**3
    if free_gb < cfg.BIG_CUDA_THRESHOLD:
        return cfg.SMALL_CUDA_BATCH
    return cfg.BIG_CUDA_PRED_BATCH if prediction else cfg.BIG_CUDA_TRAIN_BATCH


def rotate_array_to_axis(array: np.ndarray, axis: Axis = Axis.Z) -> np.ndarray:
    """:132-138 (views; self-inverse)."""
    if axis == Axis.Z:
        return array
    if axis == Axis.Y:
        return array.swapaxes(0, 1)
    if axis == Axis.X:
        return array.swapaxes(0, 2)


def one_hot_encode_array(input_array: np.ndarray, num_labels: int) -> np.ndarray:
    """:141-147"""
    flat = input_array.ravel()
    out = np.zeros((num_labels, flat.size), dtype=np.uint8)
    out[flat, np.arange(flat.size)] = 1
    return out.reshape((num_labels,) + input_array.shape)


def get_padded_dimension(dimension: int) -> int:
    """augmentations.py:30-44 -- next multiple of 32 (identity on multiples)."""
    d = cfg.IM_SIZE_DIVISOR
    return dimension if dimension % d == 0 else (dimension // d + 1) * d


def downsample_data(data, factor=2):
    """:160-162 -- block mean (NaN-aware) with edge padding by NaN, as
    skimage.measure.block_reduce(func=np.nanmean) pads... with zeros [ext];
    implemented with numpy only."""
    logging.info(f"Downsampling data by a factor of {factor}.")
    pads = [(0, (-s) % factor) for s in data.shape]
    padded = np.pad(np.asarray(data, dtype=float), pads, mode="constant", constant_values=0)
    z, y, x = (s // factor for s in padded.shape)
    blocks = padded.reshape(z, factor, y, factor, x, factor)
    return np.nanmean(blocks, axis=(1, 3, 5))


def numpy_from_tiff(path):
    """:165-176"""
    try:
        import imageio

        return imageio.volread(path)
    except ImportError:
        import cv2

        ok, pages = cv2.imreadmulti(str(path), flags=cv2.IMREAD_UNCHANGED)
        if not ok:
            raise IOError(f"could not read TIFF {path}")
        return np.stack(pages)


def numpy_from_hdf5(path, hdf5_path="/data", nexus=False):
    """:179-212"""
    import h5py as h5

    data_handle = h5.File(path, "r")
    if nexus:
        try:
            dataset = data_handle["processed/result/data"]
        except KeyError:
            logging.error("NXS file: Couldn't find data at 'processed/result/data' trying another path.")
            try:
                dataset = data_handle["entry/final_result_tomo/data"]
            except KeyError:
                logging.error("NXS file: Could not find entry at entry/final_result_tomo/data, exiting!")
                sys.exit(1)
    else:
        dataset = data_handle[hdf5_path]
    return dataset[()], dataset.chunks


def get_numpy_from_path(path: Path, internal_path: str = "/data"):
    """:215-233"""
    if path.suffix in cfg.TIFF_SUFFIXES:
        return numpy_from_tiff(path), True
    if path.suffix in cfg.HDF5_SUFFIXES:
        return numpy_from_hdf5(path, hdf5_path=internal_path, nexus=path.suffix == ".nxs")


def clip_to_uint8(data: np.ndarray, data_mean: float, st_dev_factor: float, cuda_device=None) -> np.ndarray:
    """:243-287 -- clip to mean +- k sigma, rescale to 0..255, truncate to uint8.
    The statistics (`np.nanstd`, clipped-voxel counts for the log) are numpy exactly as in
    the reference; the six elementwise float64 passes (:272-287) run as ONE kernel of the
    B200 engine when `cuda_device` names a GPU (bit-exact to numpy, tests/test_clip_gpu.py).
    Without a device argument this is the reference's numpy code path."""
    logging.info("Clipping data and converting to uint8.")
    logging.info("Calculating standard deviation.")
    data_st_dev = np.nanstd(data)
    logging.info(f"Std dev: {data_st_dev}. Calculating stats.")
    num_vox = data.size
    lower = data_mean - data_st_dev * st_dev_factor
    upper = data_mean + data_st_dev * st_dev_factor
    with np.errstate(invalid="ignore"):
        gt_ub = (data > upper).sum()
        lt_lb = (data < lower).sum()
    logging.info(f"Lower bound: {lower}, upper bound: {upper}")
    logging.info(f"Number of voxels above upper bound to be clipped {gt_ub} - percentage {gt_ub / num_vox * 100:.3f}%")
    logging.info(f"Number of voxels below lower bound to be clipped {lt_lb} - percentage {lt_lb / num_vox * 100:.3f}%")
    if cuda_device is not None:
        from ..engine import Engine, get_engine

        if data.dtype.name in Engine.CLIP_DTYPES and upper > lower:
            logging.info("Rescaling intensities and converting to uint8 on the GPU.")
            return get_engine(int(cuda_device)).clip_to_uint8(data, float(data_mean), float(lower), float(upper))
    if np.isnan(data).any():
        logging.info("Replacing NaN values.")
        data = np.nan_to_num(data, copy=False, nan=data_mean)
    logging.info("Rescaling intensities.")
    if np.issubdtype(data.dtype, np.integer):
        data = data.astype(float)
    data = np.clip(data, lower, upper, out=data)
    data = np.subtract(data, lower, out=data)
    data = np.divide(data, (upper - lower), out=data)
    data = np.clip(data, 0.0, 1.0, out=data)
    logging.info("Converting to uint8.")
    data = np.multiply(data, 255, out=data)
    return data.astype(np.uint8)


def save_data_to_hdf5(data, file_path, internal_path="/data", chunking=True):
    """:351-356"""
    import h5py as h5

    logging.info(f"Saving data of shape {data.shape} to {file_path}.")
    with h5.File(file_path, "w") as f:
        f.create_dataset(internal_path, data=data, chunks=chunking, compression=cfg.HDF5_COMPRESSION)
