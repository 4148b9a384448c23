from volume_segmantics_b200.host.arg_parsing import CheckExt, get_2d_prediction_parser  # noqa: F401
