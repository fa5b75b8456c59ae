"""Image read / resize / write with the semantics the reference gets from scipy 1.2.2's `scipy.misc`
(edgegan/utils/utils.py:41-54 get_image / save_images, :126-164 imread / imsave / center_crop / transform /
inverse_transform, :63-87 merge).  scipy.misc's image functions were removed from scipy long ago (and scipy 1.2.2
cannot be installed here), so they are restated on Pillow, which is what they wrapped:

  imread(path)            np.array(PIL.Image.open(path))            (palette images expanded, `flatten` -> mode 'F')
  imresize(a, (h, w))     toimage(a) -> PIL resize BILINEAR -> array
  toimage(a)              uint8 input is taken as is; ANY OTHER dtype is first min-max stretched to 0..255
                          (`bytescale`).  The reference casts every image to float right after reading it
                          (utils.py:128-130), so its resize always includes that contrast stretch -- a quirk the
                          loader must reproduce: an image whose values span 30..200 reaches the network spanning -1..1.
  imsave(path, a)         toimage(a, channel_axis=2).save(path): the same stretch over the whole merged sheet.
"""
from __future__ import annotations

import os

import numpy as np
from PIL import Image


def makedirs(path):
    """utils.py:14-22."""
    if not os.path.exists(path):
        os.makedirs(path)


def pathsplit(path):
    """Split a path into all of its components (used by EdgeGAN.test to find the class directory)."""
    parts = []
    path = os.path.normpath(path)
    while True:
        head, tail = os.path.split(path)
        if tail:
            parts.append(tail)
        elif head:
            parts.append(head)
            break
        if not head or head == path:
            break
        path = head
    return parts[::-1]


def image_manifold_size(num_images):
    """utils.py:29-33."""
    h = int(np.floor(np.sqrt(num_images)))
    w = int(np.ceil(np.sqrt(num_images)))
    assert h * w == num_images
    return h, w


def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    """scipy.misc.bytescale: uint8 passes through; everything else is stretched from [cmin, cmax] (default: the data's
    own min / max) to [low, high] and rounded half up."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if high > 255 or low < 0 or high < low:
        raise ValueError("`high` / `low` must satisfy 0 <= low <= high <= 255")
    cmin = data.min() if cmin is None else cmin
    cmax = data.max() if cmax is None else cmax
    cscale = cmax - cmin
    if cscale < 0:
        raise ValueError("`cmax` should be larger than `cmin`")
    if cscale == 0:
        cscale = 1
    scale = float(high - low) / cscale
    out = (data - cmin) * scale + low
    return (out.clip(low, high) + 0.5).astype(np.uint8)


def _toimage(arr):
    """scipy.misc.toimage for the cases the reference hits: 2-D -> 'L', H x W x {3, 4} -> 'RGB' / 'RGBA'."""
    data = np.asarray(arr)
    if np.iscomplexobj(data):
        raise ValueError("cannot convert a complex-valued array")
    if data.ndim == 2:
        return Image.fromarray(bytescale(data))
    if data.ndim == 3 and data.shape[2] in (3, 4):
        return Image.fromarray(np.ascontiguousarray(bytescale(data)))       # 3 channels -> 'RGB', 4 -> 'RGBA'
    raise ValueError("'arr' does not have a suitable array shape for any mode")


def imread(path, grayscale=False):
    """utils.py:126-130: scipy.misc.imread(path[, flatten=True]).astype(float)."""
    im = Image.open(path)
    if grayscale:
        im = im.convert("F")
    elif im.mode == "P":                         # scipy.misc.fromimage expands palette images
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    elif im.mode == "1":
        im = im.convert("L")
    return np.array(im).astype(np.float64)


def imresize(arr, size, interp="bilinear"):
    """scipy.misc.imresize(arr, [height, width]) -> uint8 array."""
    func = {"nearest": Image.NEAREST, "lanczos": Image.LANCZOS, "bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC,
            "cubic": Image.BICUBIC}[interp]
    im = _toimage(arr)
    return np.array(im.resize((int(size[1]), int(size[0])), resample=func))


def center_crop(x, crop_h, crop_w, resize_h=64, resize_w=64):
    """utils.py:138-145."""
    if crop_w is None:
        crop_w = crop_h
    h, w = x.shape[:2]
    j = int(round((h - crop_h) / 2.))
    i = int(round((w - crop_w) / 2.))
    return imresize(x[j:j + crop_h, i:i + crop_w], [resize_h, resize_w])


def transform(image, input_height, input_width, resize_height=64, resize_width=64, crop=True):
    """utils.py:148-160: (optional centre crop,) resize, then bytes -> [-1, 1]."""
    if crop:
        out = center_crop(image, input_height, input_width, resize_height, resize_width)
    else:
        out = imresize(image, [resize_height, resize_width])
    return np.array(out) / 127.5 - 1.


def inverse_transform(images):
    """utils.py:163-164."""
    return (images + 1.) / 2.


def get_image(image_path, input_height, input_width, resize_height=64, resize_width=64, crop=True, grayscale=False):
    """utils.py:41-50."""
    return transform(imread(image_path, grayscale), input_height, input_width, resize_height, resize_width, crop)


_TO_UNIT = (np.arange(256) / 127.5 - 1.).astype(np.float32)      # byte -> [-1, 1]: float64 arithmetic, then float32


def get_image_bytes(image_path, input_height, input_width, resize_height=64, resize_width=64, crop=True,
                    grayscale=False):
    """The image as `transform` sees it just before its last line: uint8 [h, w, c] after the (float) min-max stretch
    and the resize; `_TO_UNIT[result]` is then bit-identical to `get_image(...).astype(np.float32)`
    (tests/test_data_cpu.py).  The file's pixels are bytes, so the stretch is a 256-entry table built with the same
    float64 expression and the float64 round trip of the whole image is skipped.  Grayscale / non-byte images return
    the float64 result of `get_image` instead."""
    if grayscale:
        return get_image(image_path, input_height, input_width, resize_height, resize_width, crop, True)
    im = Image.open(image_path)
    if im.mode == "P":
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    elif im.mode == "1":
        im = im.convert("L")
    a = np.asarray(im)
    if a.dtype != np.uint8:
        return get_image(image_path, input_height, input_width, resize_height, resize_width, crop, False)
    if crop:
        crop_w = input_height if input_width is None else input_width
        h, w = a.shape[:2]
        j = int(round((h - input_height) / 2.))
        i = int(round((w - crop_w) / 2.))
        a = a[j:j + input_height, i:i + crop_w]
    lo, hi = int(a.min()), int(a.max())
    lut = bytescale(np.arange(256, dtype=np.float64), cmin=float(lo), cmax=float(hi))
    st = lut[a]
    if (st.shape[0], st.shape[1]) != (int(resize_height), int(resize_width)):
        st = np.array(Image.fromarray(np.ascontiguousarray(st)).resize((int(resize_width), int(resize_height)),
                                                                       resample=Image.BILINEAR))
    return st


def get_image_fast(image_path, input_height, input_width, resize_height=64, resize_width=64, crop=True,
                   grayscale=False):
    """`get_image(...)` as float32, for the loader's hot loop (see get_image_bytes)."""
    st = get_image_bytes(image_path, input_height, input_width, resize_height, resize_width, crop, grayscale)
    return _TO_UNIT[st] if st.dtype == np.uint8 else st.astype(np.float32)


def merge_images(images, size):
    """utils.py:63-64."""
    return inverse_transform(images)


def merge(images, size):
    """utils.py:67-87: tile a batch [N, h, w, c] into a (size[0] x size[1]) sheet, row-major."""
    images = np.asarray(images)
    h, w = images.shape[1], images.shape[2]
    c = images.shape[3]
    if c not in (1, 3, 4):
        raise ValueError("in merge(images,size) images parameter must have dimensions: HxW or HxWx3 or HxWx4")
    sheet = np.zeros((h * size[0], w * size[1], c))
    for idx, image in enumerate(images):
        col, row = idx % size[1], idx // size[1]
        sheet[row * h:(row + 1) * h, col * w:(col + 1) * w, :] = image
    return sheet[:, :, 0] if c == 1 else sheet


def imsave(images, size, path):
    """utils.py:133-135: scipy.misc.imsave(path, squeeze(merge(images, size))) -- min-max stretched to 0..255 on save."""
    image = np.squeeze(merge(images, size))
    _toimage(image).save(path)


def save_images(images, size, image_path):
    """utils.py:53-54."""
    return imsave(inverse_transform(images), size, image_path)
