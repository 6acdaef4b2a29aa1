from deadtrees_b200.utils.timer import record_execution_time  # noqa: F401
