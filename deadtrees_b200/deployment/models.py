"""Response metadata of the REST entry point - mirrors ``deadtrees/deployment/models.py:6-14``."""
import json

from pydantic import BaseModel


class PredictionStats(BaseModel):
    fraction: float
    model_name: str
    model_type: str
    elapsed: float


def predictionstats_to_str(stats: PredictionStats):
    """header-safe dict: every float as a string (``models.py:13-14``)."""
    dump = stats.model_dump_json() if hasattr(stats, "model_dump_json") else stats.json()
    return json.loads(dump, parse_float=str)
