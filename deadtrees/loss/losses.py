from deadtrees_b200.loss.losses import (EPS, BoundaryLoss, DiceLoss, FocalLoss, SurfaceLoss, class2one_hot,  # noqa: F401
                                        one_hot2dist)
