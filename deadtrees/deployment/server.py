from deadtrees_b200.deployment.server import MODEL, ModelTypes, app, create_app, segment_image  # noqa: F401
