from deadtrees_b200.data.deadtreedata import DeadtreeDatasetConfig, val_transform  # noqa: F401
