"""Quality / Axis / ModelType enums of the reference
(volume_segmantics/utilities/base_data_utils.py:21-50).

``.pytorch`` model files pickle ``ModelType`` by its qualified name
``volume_segmantics.utilities.base_data_utils.ModelType`` (written by
early_stopping.py:55-62), so the classes advertise that module and the
``volume_segmantics`` compatibility package re-exports these very objects.
"""
from enum import Enum

_REF_MODULE = "volume_segmantics.utilities.base_data_utils"


class Quality(Enum):
    """Number of slicing directions: 1 axis, 3 axes, 3 axes x 4 rotations."""
    LOW = 1
    MEDIUM = 3
    HIGH = 12


class Axis(Enum):
    Z = 0
    Y = 1
    X = 2
    ALL = 4


class ModelType(Enum):
    U_NET = 1
    U_NET_PLUS_PLUS = 2
    FPN = 3
    DEEPLABV3 = 4
    DEEPLABV3_PLUS = 5
    MA_NET = 6
    LINKNET = 7
    PAN = 8


for _cls in (Quality, Axis, ModelType):
    _cls.__module__ = _REF_MODULE
