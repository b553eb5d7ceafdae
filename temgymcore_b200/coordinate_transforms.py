"""3x3 homogeneous pixel<->metre transforms in (y, x, 1) order.

Host-side restatement of reference ``src/temgym_core/coordinate_transforms.py:17-101``
(six tiny 3x3 matrices; numpy fp64, same composition order).  The per-point application
of the matrix (``apply_transformation``, coordinate_transforms.py:104-140) runs on the GPU
(``tg_metres_to_pixels``, csrc/trace.cu).
"""
import numpy as np


def _rotate(radians):
    # y, x instead of x, y; y-axis flipped convention (coordinate_transforms.py:17-30)
    c, s = np.cos(radians), np.sin(radians)
    return np.array([(c, s, 0.0), (-s, c, 0.0), (0.0, 0.0, 1.0)])


def _rotate_with_deg_to_rad(degrees):
    return _rotate(np.pi / 180 * degrees)


def _identity():
    return np.eye(3)


def _scale(pixel_size_yx):
    return np.array([(pixel_size_yx[0], 0.0, 0.0), (0.0, pixel_size_yx[1], 0.0), (0.0, 0.0, 1.0)])


def _shift(centre_yx):
    return np.array([(1.0, 0.0, centre_yx[0]), (0.0, 1.0, centre_yx[1]), (0.0, 0.0, 1.0)])


def _flip_y():
    return np.array([(-1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)])


def pixels_to_metres_transform(centre, pixel_size, shape, flip_y=False, rotation=0.0):
    """coordinate_transforms.py:50-101; composition order of lines 92-99."""
    flip_transform = _flip_y() if flip_y else _identity()
    shape = np.array(shape)
    return (
        flip_transform
        @ _flip_y()
        @ _rotate_with_deg_to_rad(rotation)
        @ _shift(centre)
        @ _scale(pixel_size)
        @ _shift(-(shape - 1) / 2.0)
    )


def apply_transformation(y, x, transformation):
    """``T @ [y, x, 1]`` on the GPU (coordinate_transforms.py:104-140)."""
    from .grid import _affine_apply
    return _affine_apply(y, x, np.asarray(transformation, dtype=np.float64), as_float=True)
