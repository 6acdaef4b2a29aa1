from deadtrees_b200.network.segmodel import *  # noqa: F401,F403
from deadtrees_b200.network.segmodel import SemSegment, create_combined_batch, initialize_weights  # noqa: F401
