"""Small containers shared by the result objects (reference: utils/helpers.py:9-14)."""
from collections import defaultdict


def nested_dict():
    """Arbitrarily deep auto-vivifying dictionary."""
    return defaultdict(nested_dict)
