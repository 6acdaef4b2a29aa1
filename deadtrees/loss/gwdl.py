from deadtrees_b200.loss.gwdl import GeneralizedWassersteinDiceLoss  # noqa: F401
