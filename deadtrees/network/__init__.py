from deadtrees_b200.network import SemSegment, Unet  # noqa: F401
