from volume_segmantics_b200.host.base_data_manager import BaseDataManager  # noqa: F401
