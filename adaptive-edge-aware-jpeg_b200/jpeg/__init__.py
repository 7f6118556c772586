from .jpeg import Jpeg, JpegCompressionSettings

__all__ = ["Jpeg", "JpegCompressionSettings"]
