"""Shim: the reference's import paths (``deadtrees.*``) served by ``deadtrees_b200``."""
