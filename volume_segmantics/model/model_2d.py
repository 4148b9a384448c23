from volume_segmantics_b200.host.model_2d import create_model_from_file, create_model_on_device  # noqa: F401
