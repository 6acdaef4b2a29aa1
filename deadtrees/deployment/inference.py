from deadtrees_b200.deployment.inference import Inference, PyTorchInference, PyTorchEnsembleInference, MosaicInference  # noqa: F401
