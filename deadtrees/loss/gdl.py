from deadtrees_b200.loss.gdl import GeneralizedDiceLoss  # noqa: F401
