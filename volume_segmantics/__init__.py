"""Import-path compatibility package: ``volume_segmantics.*`` names used by the
reference's callers (SuRVoS2, the CLI, pickled ``.pytorch`` files) resolve to the
B200 implementation in ``volume_segmantics_b200``.  Prediction path only."""
