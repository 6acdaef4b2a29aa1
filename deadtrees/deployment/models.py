from deadtrees_b200.deployment.models import PredictionStats, predictionstats_to_str  # noqa: F401
