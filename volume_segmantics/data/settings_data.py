from volume_segmantics_b200.host.settings_data import get_settings_data  # noqa: F401
