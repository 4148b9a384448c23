"""volseg-b200: B200 (sm_100a) engine for the prediction hot path of
DiamondLightSource/volume-segmantics.  See DESIGN.md."""
__version__ = "0.1.0"
