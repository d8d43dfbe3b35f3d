"""Import stub so the reference's PerceptualSimilarity package can be imported (gen_golden.py only)."""
from . import measure, transform, color  # noqa
