from deadtrees_b200.data.deadtreedata import (DeadtreeDatasetConfig, BatchTrainTransform, train_transform, transform,  # noqa: F401
                                               val_transform)
