from .conversion import apply_normalization, convert, get_color_spaces

__all__ = ["apply_normalization", "convert", "get_color_spaces"]
