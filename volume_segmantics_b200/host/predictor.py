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


def _as_engine_volume(data_vol: np.ndarray) -> np.ndarray:
    """The reference slices whatever array it is given: integer slices of ANY bit depth are cast to
    float32 and divided by 255, float32 slices are fed as they are (datasets.py:129-135).  The engine's
    slicer does the same for uint8 / int8 / uint16 / int16 / int32 / float32 volumes.  Wider integers
    are narrowed to int32 when their values fit (cv2.copyMakeBorder, which pads the reference's slices,
    converts int64 to int32 itself).  float64 / float16 volumes fail in the reference as well -- the
    network receives a double / half batch for float32 weights (vol_seg_2d_predictor.py:44) -- and are
    refused here with the remedy."""
    vol = np.asarray(data_vol)
    if vol.ndim != 3:
        raise ValueError(f"expected a 3-D volume, got shape {vol.shape}")
    if vol.dtype.name in Engine.VOLUME_DTYPES:
        return vol if vol.flags.c_contiguous else np.ascontiguousarray(vol)
    if np.issubdtype(vol.dtype, np.integer):
        lo, hi = int(vol.min()), int(vol.max())
        if lo >= -(2**31) and hi < 2**31:
            return np.ascontiguousarray(vol.astype(np.int32))
        raise ValueError(f"integer volume with range [{lo}, {hi}] does not fit int32 (cv2 pads int32 at most)")
    if vol.dtype == np.bool_:
        return np.ascontiguousarray(vol.astype(np.uint8))
    raise TypeError(
        f"volume dtype {vol.dtype}: the reference feeds float volumes to the float32 network unchanged, which only "
        "works for float32; convert the volume to float32 or enable clip_data (2d_model_predict_settings.yaml:4)"
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
        # additive setting (absent => the reference's single device): several GPUs of this host
        devices = getattr(settings, "cuda_devices", None)
        self._devices = [int(d) for d in devices] if devices and len(list(devices)) > 1 else None
        self._group = None

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
        # re-lowers when .model was replaced or its weights changed since they were lowered
        self.engine.ensure_model(self.model)
        self.engine.set_vote_mode(False)
        if isinstance(data_vol, np.ndarray) and self.engine.holds(data_vol):
            # the array BaseDataManager's GPU pre-processing returned: its content is already resident
            self.engine.reset_for(data_vol.shape)
            return data_vol
        vol = _as_engine_volume(data_vol)
        self.engine.set_volume(vol)
        return vol

    def _get_model_from_trainer(self, trainer):
        # :28-29 ; the new module's weights are folded on the next prediction
        self.model = trainer.model
        if self._engine is not None:
            self._engine._plan_model = None

    @property
    def group(self):
        if self._group is None:
            from ..multi import LocalGroup

            self._group = LocalGroup(self._devices)
        return self._group

    def _run(self, data_vol, dir_mask: int, output_probs: bool = True):
        if self._devices is not None:
            vol = _as_engine_volume(data_vol)
            if vol.dtype == np.uint8:
                if not isinstance(self.model, B200SegmentationModel):
                    raise TypeError("VolSeg2dPredictor.model must be a B200SegmentationModel")
                return self.group.predict(self.model, vol, dir_mask, want_probs=output_probs)
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
