"""``model-predict-2d <model> <data> [--data_dir DIR]``
(volume_segmantics/scripts/predict_2d_model.py:16-38)."""
import logging
import warnings
from datetime import date
from pathlib import Path

from . import constants as cfg
from .arg_parsing import get_2d_prediction_parser
from .manager import VolSeg2DPredictionManager
from .settings_data import get_settings_data

warnings.filterwarnings("ignore", category=UserWarning)


def create_output_path(root_path, data_vol_path):
    return Path(root_path, f"{date.today()}_{data_vol_path.stem}_2d_model_vol_pred.h5")


def main():
    logging.basicConfig(level=logging.INFO, format=cfg.LOGGING_FMT, datefmt=cfg.LOGGING_DATE_FMT)
    args = get_2d_prediction_parser().parse_args()
    root_path = Path(getattr(args, cfg.DATA_DIR_ARG)).resolve()
    settings = get_settings_data(Path(root_path, cfg.SETTINGS_DIR, cfg.PREDICTION_SETTINGS_FN))
    data_vol_path = Path(getattr(args, cfg.PREDICT_DATA_ARG))
    manager = VolSeg2DPredictionManager(getattr(args, cfg.MODEL_PTH_ARG), data_vol_path, settings)
    manager.predict_volume_to_path(create_output_path(root_path, data_vol_path))


if __name__ == "__main__":
    main()
