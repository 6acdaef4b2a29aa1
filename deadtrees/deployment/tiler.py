from deadtrees_b200.deployment.tiler import Tiler, TileInfo, divisible_without_remainder, inspect_tile, inspect_array  # noqa: F401
