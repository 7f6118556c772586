from .evaluation_metrics import EvaluationMetrics
from .image import Image

__all__ = ["EvaluationMetrics", "Image"]
