from volume_segmantics_b200.host.utils import get_padded_dimension  # noqa: F401
