"""Drop-in ``VolSeg2dPredictor`` over the B200 engine.

Same constructor, attributes, method names, argument meaning, return dtypes and
shapes as volume_segmantics/model/operations/vol_seg_2d_predictor.py:16-136;
each method cites the lines it replaces.  The slice loop, network, softmax,
crop, rotation and merges all run inside libvsb200 (one call per request);
this class only converts the input to the uint8 volume the reference would have
sliced, and picks the direction set.
"""
from __future__ import annotations

import logging
from pathlib import Path
from types import SimpleNamespace

import numpy as np

from ..engine import Engine, get_engine
from ..plan import B200SegmentationModel
from .enums import Axis
from .model_2d import create_model_from_file

DIRS_3 = 0b111
DIRS_12 = (1 << 12) - 1


def _as_uint8_volume(data_vol: np.ndarray) -> np.ndarray:
    """The reference divides any integer slice by 255 and feeds floats as they
    are (datasets.py:129-135).  The engine ingests uint8 (what clip_to_uint8
    produces, base_data_utils.py:243-287); integer volumes already inside
    0..255 convert losslessly, anything else is refused loudly."""
    vol = np.asarray(data_vol)
    if vol.ndim != 3:
        raise ValueError(f"expected a 3-D volume, got shape {vol.shape}")
    if vol.dtype == np.uint8:
        return np.ascontiguousarray(vol)
    if np.issubdtype(vol.dtype, np.integer) or vol.dtype == np.bool_:
        lo, hi = int(vol.min()), int(vol.max())
        if lo >= 0 and hi <= 255:
            return np.ascontiguousarray(vol.astype(np.uint8))
        raise NotImplementedError(
            f"integer volume with range [{lo}, {hi}] outside 0..255: enable clip_data "
            "(the B200 slicer ingests uint8 volumes only)"
        )
    raise NotImplementedError(
        f"volume dtype {vol.dtype} is not supported by the B200 slicer: enable clip_data so "
        "the data is rescaled to uint8 first (reference default, 2d_model_predict_settings.yaml:4)"
    )


class VolSeg2dPredictor:
    """Performs 2-D network prediction over a 3-D volume. Does not touch disk."""

    def __init__(self, model_file_path: str, settings: SimpleNamespace) -> None:
        # vol_seg_2d_predictor.py:19-26
        self.model_file_path = Path(model_file_path)
        self.settings = settings
        self.model_device_num = int(settings.cuda_device)
        self.model, self.num_labels, self.label_codes = create_model_from_file(
            self.model_file_path, True, self.model_device_num
        )
        self._engine = None

    # -- engine plumbing -------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = get_engine(self.model_device_num)
        return self._engine

    def _prepare(self, data_vol) -> np.ndarray:
        if not isinstance(self.model, B200SegmentationModel):
            raise TypeError(
                "VolSeg2dPredictor.model must be a B200SegmentationModel; a foreign nn.Module "
                "cannot run on the B200 engine and there is no PyTorch fallback"
            )
        self.model._engine = self.engine
        self.engine.ensure_model(self.model)  # re-lowers if .model was replaced
        vol = _as_uint8_volume(data_vol)
        self.engine.set_vote_mode(False)
        self.engine.set_volume(vol)
        return vol

    def _get_model_from_trainer(self, trainer):
        # :28-29 ; weights are re-folded on the next prediction
        self.model = trainer.model

    def _run(self, data_vol, dir_mask: int, output_probs: bool = True):
        self._prepare(data_vol)
        self.engine.predict(dir_mask, skip_duplicates=True)
        return self.engine.fetch(want_probs=output_probs)

    # -- :31-65 ------------------------------------------------------------------
    def _predict_single_axis(self, data_vol, output_probs=True, axis=Axis.Z):
        if axis not in (Axis.Z, Axis.Y, Axis.X):
            raise ValueError(f"prediction axis must be Z, Y or X, got {axis}")
        shape = tuple(np.asarray(data_vol).swapaxes(0, axis.value).shape) if axis != Axis.Z else data_vol.shape
        logging.info(f"Predicting segmentation for volume of shape {shape}.")
        labels, probs = self._run(data_vol, 1 << axis.value, output_probs)
        return labels, probs

    # -- :67-88 (+ :90-98 merges, fused) -------------------------------------------
    def _predict_3_ways_max_probs(self, data_vol):
        logging.info("Predicting YX, ZX and ZY slices; merging by maximum probability on the GPU.")
        return self._run(data_vol, DIRS_3)

    # -- :100-116 ------------------------------------------------------------------
    def _predict_12_ways_max_probs(self, data_vol):
        logging.info("Predicting 3 axes x 4 rotations; merging by maximum probability on the GPU.")
        return self._run(data_vol, DIRS_12)

    # -- :90-98 : kept for callers that merge their own containers -------------------
    def _merge_vols_in_mem(self, prob_container, label_container):
        max_prob_idx = np.argmax(prob_container, axis=0)[np.newaxis]
        prob_container[0] = np.squeeze(np.take_along_axis(prob_container, max_prob_idx, axis=0), axis=0)
        label_container[0] = np.squeeze(np.take_along_axis(label_container, max_prob_idx, axis=0), axis=0)

    # -- :118-136 one-hot votes ------------------------------------------------------
    def _votes(self, data_vol, dir_mask: int) -> np.ndarray:
        self._prepare(data_vol)
        self.engine.set_vote_mode(True)
        self.engine.reset()
        self.engine.predict(dir_mask, skip_duplicates=False)
        votes = self.engine.fetch_votes()
        self.engine.set_vote_mode(False)
        return votes

    def _predict_single_axis_to_one_hot(self, data_vol, axis=Axis.Z):
        return self._votes(data_vol, 1 << axis.value)

    def _predict_3_ways_one_hot(self, data_vol):
        return self._votes(data_vol, DIRS_3)

    def _predict_12_ways_one_hot(self, data_vol):
        return self._votes(data_vol, DIRS_12)

    # older public names that appear in the reference's generated docs
    predict_single_axis = _predict_single_axis
    predict_3_ways_max_probs = _predict_3_ways_max_probs
    predict_12_ways_max_probs = _predict_12_ways_max_probs
    merge_vols_in_mem = _merge_vols_in_mem
    predict_single_axis_to_one_hot = _predict_single_axis_to_one_hot
    predict_3_ways_one_hot = _predict_3_ways_one_hot
    predict_12_ways_one_hot = _predict_12_ways_one_hot
