"""Volume loading + pre-processing in front of the hot path
(volume_segmantics/data/base_data_manager.py:10-42)."""
import logging
from pathlib import Path
from types import SimpleNamespace
from typing import Union

import numpy as np

from . import utils


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


class BaseDataManager:
    def __init__(self, data_vol: Union[Path, str, np.ndarray], settings: SimpleNamespace) -> None:
        self.data_vol_shape = None
        self.data_mean = None
        self.data_vol_path = utils.setup_path_if_exists(data_vol)
        self.settings = settings
        self.st_dev_factor = settings.st_dev_factor
        self.downsample = settings.downsample
        if self.data_vol_path is not None:
            self.data_vol, self.input_data_chunking = utils.get_numpy_from_path(
                self.data_vol_path, internal_path=settings.data_hdf5_path
            )
        elif isinstance(data_vol, np.ndarray):
            self.data_vol = data_vol
            self.input_data_chunking = True
        self._preprocess_data()

    def _preprocess_data(self):
        # base_data_manager.py:29-42
        if self.downsample:
            self.data_vol = utils.downsample_data(self.data_vol)
        self.data_vol_shape = self.data_vol.shape
        device = getattr(self.settings, "cuda_device", None) if _cuda_available() else None
        if device is not None and self.settings.clip_data and self._preprocess_on_gpu(int(device)):
            return
        logging.info("Calculating mean of data...")
        self.data_mean = np.nanmean(self.data_vol)
        logging.info(f"Mean value: {self.data_mean}")
        if self.settings.clip_data:
            # elementwise part on the GPU the prediction will use (None -> numpy, as the reference)
            self.data_vol = utils.clip_to_uint8(self.data_vol, self.data_mean, self.st_dev_factor, cuda_device=device)
        if np.isnan(self.data_vol).any():
            logging.info("Replacing NaN values.")
            self.data_vol = np.nan_to_num(self.data_vol, copy=False)

    def _preprocess_on_gpu(self, device: int) -> bool:
        """nanmean, nanstd, the two clipped-voxel counts and the clip / rescale / quantise pass of
        base_data_manager.py:33-39 + base_data_utils.py:243-287 on the GPU: ONE upload of the raw
        volume; the uint8 result stays in HBM as the engine's volume (the prediction that follows does
        not upload it again) and comes back once as a read-only array for ``self.data_vol``.
        The statistics are float64 reductions with a fixed order: equal to numpy's pairwise sums to
        rounding (not bit for bit); given the statistics the elementwise pass is bit-exact."""
        from ..engine import Engine, get_engine

        data = self.data_vol
        if not isinstance(data, np.ndarray) or data.ndim != 3 or data.dtype.name not in Engine.CLIP_DTYPES:
            return False
        eng = get_engine(device)
        eng.raw_upload(data)
        logging.info("Calculating mean of data...")
        count, mean, st_dev, nans = eng.raw_moments()
        # numpy hands back float32 scalars for float32 data and float64 otherwise; the bounds are then
        # computed in that type (base_data_utils.py:259-260)
        ftype = np.float32 if data.dtype == np.float32 else np.float64
        self.data_mean = ftype(mean)
        logging.info(f"Mean value: {self.data_mean}")
        logging.info("Clipping data and converting to uint8.")
        data_st_dev = ftype(st_dev)
        logging.info(f"Std dev: {data_st_dev}. Calculating stats.")
        lower = self.data_mean - (data_st_dev * self.st_dev_factor)
        upper = self.data_mean + (data_st_dev * self.st_dev_factor)
        if not (count > 0 and upper > lower):  # constant or all-NaN volume: leave it to numpy, as the reference
            eng.raw_release()
            return False
        logging.info(f"Lower bound: {lower}, upper bound: {upper}")
        out, gt_ub, lt_lb = eng.raw_clip_to_volume(float(self.data_mean), float(lower), float(upper), data.shape)
        num_vox = data.size
        logging.info(f"Number of voxels above upper bound to be clipped {gt_ub} - percentage {gt_ub / num_vox * 100:.3f}%")
        logging.info(f"Number of voxels below lower bound to be clipped {lt_lb} - percentage {lt_lb / num_vox * 100:.3f}%")
        if nans:
            logging.info("Replacing NaN values.")
        logging.info("Rescaling intensities.")
        logging.info("Converting to uint8.")
        self.data_vol = out
        return True
