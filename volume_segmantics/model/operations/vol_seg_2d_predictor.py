from volume_segmantics_b200.host.predictor import VolSeg2dPredictor  # noqa: F401
