"""jpeg.utils drop-in (src/jpeg/utils.py:24-41)."""


def largest_power_of_2(n: int) -> int:
    """n for n <= 2, otherwise the largest power of two STRICTLY below n -- the reference's
    behaviour (its docstring says "less than or equal", its code computes 2**floor(log2(n-1)))."""
    if n <= 0:
        raise ValueError("n must be positive.")
    if n <= 2:
        return n
    return 1 << ((n - 1).bit_length() - 1)
