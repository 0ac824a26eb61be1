"""Minimal NIfTI-1 (.nii / .nii.gz) writer and reader for the posterior maps of ``save_predictions``
(reference model.py:792-802 uses nibabel, which is not installed here).  Single-file format, float32 / float64 /
int16 / uint8 data, identity affine unless one is given -- what ``nib.Nifti1Image(array, None)`` produces."""
from __future__ import annotations

import gzip
import struct

import numpy as np

_DTYPES = {np.dtype('uint8'): (2, 8), np.dtype('int16'): (4, 16), np.dtype('int32'): (8, 32),
           np.dtype('float32'): (16, 32), np.dtype('float64'): (64, 64)}
_CODES = {code: dt for dt, (code, _) in _DTYPES.items()}


def save_nifti(array, path, affine=None):
    """Write ``array`` (up to 7-D, first three axes spatial) to ``path`` ('.nii' or '.nii.gz')."""
    a = np.asarray(array)
    if a.dtype not in _DTYPES:
        a = a.astype(np.float32)
    if not 1 <= a.ndim <= 7:
        raise ValueError('NIfTI-1 stores 1 to 7 dimensions, got %d' % a.ndim)
    code, bitpix = _DTYPES[a.dtype]
    aff = np.eye(4, dtype=np.float32) if affine is None else np.asarray(affine, dtype=np.float32).reshape(4, 4)
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    struct.pack_into('<8h', hdr, 40, a.ndim, *(list(a.shape) + [1] * (7 - a.ndim)))
    struct.pack_into('<h', hdr, 70, code)
    struct.pack_into('<h', hdr, 72, bitpix)
    struct.pack_into('<8f', hdr, 76, 1.0, *([1.0] * 7))
    struct.pack_into('<f', hdr, 108, 352.0)                      # vox_offset
    struct.pack_into('<2f', hdr, 112, 1.0, 0.0)                  # scl_slope, scl_inter
    struct.pack_into('<2h', hdr, 252, 0, 2 if affine is not None else 0)    # qform_code, sform_code
    for i in range(3):
        struct.pack_into('<4f', hdr, 280 + 16 * i, *aff[i])
    hdr[344:348] = b'n+1\x00'
    payload = bytes(hdr) + b'\x00' * 4 + np.asfortranarray(a).tobytes(order='F')
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'wb') as f:
        f.write(payload)


def load_nifti(path):
    """Return (array, affine) of a single-file NIfTI-1 image written by save_nifti (or any little-endian one)."""
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'rb') as f:
        raw = f.read()
    if struct.unpack_from('<i', raw, 0)[0] != 348 or raw[344:347] != b'n+1':
        raise ValueError('%s is not a little-endian single-file NIfTI-1 image' % path)
    dim = struct.unpack_from('<8h', raw, 40)
    shape = tuple(dim[1:1 + dim[0]])
    code = struct.unpack_from('<h', raw, 70)[0]
    off = int(struct.unpack_from('<f', raw, 108)[0])
    slope, inter = struct.unpack_from('<2f', raw, 112)
    dt = _CODES[code]
    data = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=off).reshape(shape, order='F')
    if slope not in (0.0, 1.0) and np.isfinite(slope) or inter != 0.0:
        data = data * slope + inter
    aff = np.eye(4, dtype=np.float32)
    for i in range(3):
        aff[i] = struct.unpack_from('<4f', raw, 280 + 16 * i)
    return np.array(data), aff


def save_im_data(im_data, filename, affine=None):
    """save_im_data of the reference (model.py:792-802): [S,X,Y,Z,C] -> one [X,Y,Z,S*C] image at filename.nii.gz
    (subjects concatenated along the last axis)."""
    im = np.asarray(im_data)
    images = np.concatenate(np.split(im, im.shape[0], axis=0), axis=-1)[0]
    save_nifti(images.astype(np.float32), filename + '.nii.gz', affine)
