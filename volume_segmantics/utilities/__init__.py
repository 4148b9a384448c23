from volume_segmantics_b200.host.arg_parsing import get_2d_prediction_parser
from volume_segmantics_b200.host.enums import Quality

__all__ = ["get_2d_prediction_parser", "Quality"]
