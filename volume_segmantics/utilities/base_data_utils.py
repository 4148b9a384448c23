"""Re-exports; ``ModelType`` must be reachable here because model files pickle
it as ``volume_segmantics.utilities.base_data_utils.ModelType``."""
from volume_segmantics_b200.host.enums import Axis, ModelType, Quality  # noqa: F401
from volume_segmantics_b200.host.utils import *  # noqa: F401,F403
from volume_segmantics_b200.host.utils import (  # noqa: F401
    clip_to_uint8, create_enum_from_setting, downsample_data, get_batch_size, get_model_type,
    get_numpy_from_path, get_padded_dimension, get_prediction_axis, get_prediction_quality,
    get_training_axis, numpy_from_hdf5, numpy_from_tiff, one_hot_encode_array,
    rotate_array_to_axis, save_data_to_hdf5, setup_path_if_exists,
)
