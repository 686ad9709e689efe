from .common import GaussianHeatmapGenerator, PoseRegressionHead  # noqa: F401
