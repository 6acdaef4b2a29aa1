from deadtrees_b200.utils.data_handling import make_blocks_vectorized, unmake_blocks_vectorized  # noqa: F401
