"""Drop-in ``VolSeg2DPredictionManager``
(volume_segmantics/model/operations/vol_seg_prediction_manager.py:11-100)."""
from pathlib import Path
from types import SimpleNamespace
from typing import Union

import numpy as np

from . import utils
from .base_data_manager import BaseDataManager
from .enums import Quality
from .predictor import VolSeg2dPredictor


class VolSeg2DPredictionManager(BaseDataManager):
    """Manages prediction of a data volume to disk with a 2-D network."""

    def __init__(self, model_file_path: str, data_vol: Union[str, np.ndarray], settings: SimpleNamespace) -> None:
        super().__init__(data_vol, settings)  # :30
        self.predictor = VolSeg2dPredictor(model_file_path, settings)  # :31
        self.settings = settings

    def get_label_codes(self) -> dict:
        return self.predictor.label_codes  # :34-41

    def predict_volume_to_path(
        self, output_path: Union[Path, None], quality: Union[Quality, None] = None
    ) -> np.ndarray:
        """:43-100 -- LOW: one axis; MEDIUM: 3 axes; HIGH: 3 axes x 4 rotations,
        merged by maximum probability (or summed one-hot votes)."""
        probs = None
        one_hot = self.settings.one_hot
        preferred_axis = utils.get_prediction_axis(self.settings)
        if quality is None:
            quality = utils.get_prediction_quality(self.settings)
        p = self.predictor
        if quality == Quality.LOW:
            if one_hot:
                prediction = p._predict_single_axis_to_one_hot(self.data_vol, axis=preferred_axis)
            else:
                prediction, probs = p._predict_single_axis(self.data_vol, axis=preferred_axis)
        if quality == Quality.MEDIUM:
            if one_hot:
                prediction = p._predict_3_ways_one_hot(self.data_vol)
            else:
                prediction, probs = p._predict_3_ways_max_probs(self.data_vol)
        if quality == Quality.HIGH:
            if one_hot:
                prediction = p._predict_12_ways_one_hot(self.data_vol)
            else:
                prediction, probs = p._predict_12_ways_max_probs(self.data_vol)
        if output_path is not None:
            output_path = Path(output_path)
            utils.save_data_to_hdf5(prediction, output_path, chunking=self.input_data_chunking)
            if probs is not None and self.settings.output_probs:
                utils.save_data_to_hdf5(
                    probs, f"{output_path.parent / output_path.stem}_probs.h5", chunking=self.input_data_chunking
                )
        return prediction
