from volume_segmantics_b200.host.manager import VolSeg2DPredictionManager  # noqa: F401
