"""`edgegan/utils/data/__init__.py:1`."""
from .dataset import Dataset, DevicePrefetcher, extension_match_recursive  # noqa: F401
