from volume_segmantics_b200.host.constants import *  # noqa: F401,F403
