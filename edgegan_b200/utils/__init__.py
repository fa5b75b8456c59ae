"""Host-side image I/O of the reference (`edgegan/utils/__init__.py:1` re-exports `edgegan/utils/utils.py`)."""
from .utils import (_TO_UNIT, bytescale, center_crop, get_image, get_image_bytes, get_image_fast, image_manifold_size, imread, imresize, imsave,  # noqa: F401
                    inverse_transform, makedirs, merge, merge_images, pathsplit, save_images, transform)
