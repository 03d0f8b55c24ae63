"""Minimal zarr-v2 directory stores, written and read without the ``zarr`` package.

The reference persists neighbour info as ``xr.Dataset.to_zarr("nn_lut-<hash>.zarr")`` and reads it back with
``zarr.open(filename)[name]`` (satpy/resample/kdtree.py:125-177).  A zarr-v2 store is a directory holding ``.zgroup``,
per array a ``.zarray`` JSON document plus one raw file per chunk (``"0.0"``, ``"1.0"`` ...), and -- what xarray needs
to rebuild the dimensions -- ``_ARRAY_DIMENSIONS`` in the array's ``.zattrs``.  That layout is written here by hand
(uncompressed chunks, C order, row chunks of about 64 MiB), so a store written by this package opens with
``zarr.open`` / ``xr.open_zarr`` in a full satpy installation, and stores whose chunks are uncompressed or
zlib/gzip-compressed are read back; chunks compressed with blosc (zarr's default) need the ``zarr`` package, which is
used when importable.
"""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

_TARGET_CHUNK_BYTES = 64 << 20


def _chunk_shape(shape, itemsize):
    if not shape:
        return ()
    row_bytes = int(np.prod(shape[1:], dtype=np.int64)) * itemsize
    rows = max(1, min(shape[0], _TARGET_CHUNK_BYTES // max(1, row_bytes)))
    return (int(rows),) + tuple(int(s) for s in shape[1:])


def write_store(path, arrays, attrs=None):
    """``arrays``: name -> (dims, ndarray).  Overwrites ``path``."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, ".zgroup"), "w") as fh:
        json.dump({"zarr_format": 2}, fh)
    with open(os.path.join(path, ".zattrs"), "w") as fh:
        json.dump(dict(attrs or {}), fh)
    consolidated = {".zgroup": {"zarr_format": 2}, ".zattrs": dict(attrs or {})}
    for name, (dims, arr) in arrays.items():
        arr = np.ascontiguousarray(arr)
        adir = os.path.join(path, name)
        os.makedirs(adir, exist_ok=True)
        chunks = _chunk_shape(arr.shape, arr.dtype.itemsize)
        fill = False if arr.dtype == np.bool_ else (0 if arr.dtype.kind in "iu" else "NaN")
        zarray = {"chunks": list(chunks), "compressor": None, "dtype": arr.dtype.str, "fill_value": fill,
                  "filters": None, "order": "C", "shape": list(arr.shape), "zarr_format": 2}
        zattrs = {"_ARRAY_DIMENSIONS": list(dims)}
        with open(os.path.join(adir, ".zarray"), "w") as fh:
            json.dump(zarray, fh)
        with open(os.path.join(adir, ".zattrs"), "w") as fh:
            json.dump(zattrs, fh)
        consolidated[f"{name}/.zarray"] = zarray
        consolidated[f"{name}/.zattrs"] = zattrs
        rows = chunks[0] if chunks else 1
        tail = ".".join("0" for _ in arr.shape[1:])
        for k, r0 in enumerate(range(0, arr.shape[0] if arr.shape else 1, rows)):
            block = arr[r0:r0 + rows] if arr.shape else arr
            if block.shape[0] < rows:  # edge chunks are stored full-size (zarr v2), padded with the fill value
                pad = np.zeros((rows - block.shape[0],) + block.shape[1:], dtype=arr.dtype)
                block = np.concatenate([block, pad], axis=0)
            fname = str(k) if not tail else f"{k}.{tail}"
            with open(os.path.join(adir, fname), "wb") as fh:
                fh.write(np.ascontiguousarray(block).tobytes())
    with open(os.path.join(path, ".zmetadata"), "w") as fh:
        json.dump({"metadata": consolidated, "zarr_consolidated_format": 1}, fh)


def write_root_array(path, arr):
    """A store whose root IS the array (what ``dask.array.to_zarr(path)`` writes: satpy/modifiers/angles.py:207-213)."""
    arr = np.ascontiguousarray(arr)
    os.makedirs(path, exist_ok=True)
    chunks = _chunk_shape(arr.shape, arr.dtype.itemsize)
    fill = False if arr.dtype == np.bool_ else (0 if arr.dtype.kind in "iu" else "NaN")
    with open(os.path.join(path, ".zarray"), "w") as fh:
        json.dump({"chunks": list(chunks), "compressor": None, "dtype": arr.dtype.str, "fill_value": fill,
                   "filters": None, "order": "C", "shape": list(arr.shape), "zarr_format": 2}, fh)
    rows = chunks[0] if chunks else 1
    tail = ".".join("0" for _ in arr.shape[1:])
    for k, r0 in enumerate(range(0, arr.shape[0] if arr.shape else 1, rows)):
        block = arr[r0:r0 + rows] if arr.shape else arr
        if block.shape[0] < rows:
            block = np.concatenate([block, np.zeros((rows - block.shape[0],) + block.shape[1:], dtype=arr.dtype)], axis=0)
        with open(os.path.join(path, str(k) if not tail else f"{k}.{tail}"), "wb") as fh:
            fh.write(np.ascontiguousarray(block).tobytes())


def read_root_array(path):
    """The array at the root of a store (see ``write_root_array``)."""
    head, name = os.path.split(os.path.normpath(path))
    return read_array(head, name)


def _decode(raw, compressor):
    if compressor is None:
        return raw
    cid = compressor.get("id")
    if cid == "zlib":
        return zlib.decompress(raw)
    if cid == "gzip":
        return zlib.decompress(raw, 16 + zlib.MAX_WBITS)
    raise IOError(f"zarr chunk compressor {cid!r} needs the zarr package")


def read_array(path, name):
    """One array of a zarr-v2 store as numpy (``IOError`` when the store or the array is missing / unreadable)."""
    try:
        import zarr
        root = zarr.open(path, mode="r")
        return np.array(root[name] if name else root)
    except ImportError:
        pass
    except (ValueError, KeyError) as err:  # the reference maps these to IOError too (kdtree.py:154-155)
        raise IOError(str(err)) from err
    adir = os.path.join(path, name)
    try:
        with open(os.path.join(adir, ".zarray")) as fh:
            meta = json.load(fh)
    except OSError as err:
        raise IOError(f"no zarr array {name!r} in {path}") from err
    if meta.get("zarr_format") != 2 or meta.get("filters"):
        raise IOError(f"{adir}: only unfiltered zarr-v2 arrays can be read without the zarr package")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dtype = np.dtype(meta["dtype"])
    order = meta.get("order", "C")
    out = np.empty(shape, dtype=dtype)
    sep = meta.get("dimension_separator", ".")
    grid = [range(-(-s // c)) for s, c in zip(shape, chunks)]
    for idx in np.ndindex(*[len(g) for g in grid]):
        fname = os.path.join(adir, sep.join(str(i) for i in idx))
        sl = tuple(slice(i * c, min(s, (i + 1) * c)) for i, c, s in zip(idx, chunks, shape))
        if not os.path.exists(fname):
            fv = meta.get("fill_value")
            out[sl] = 0 if fv in (None, "NaN") and dtype.kind != "f" else (np.nan if fv == "NaN" else fv)
            continue
        with open(fname, "rb") as fh:
            raw = _decode(fh.read(), meta.get("compressor"))
        block = np.frombuffer(raw, dtype=dtype).reshape(chunks, order=order)
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out


def array_dims(path, name):
    try:
        with open(os.path.join(path, name, ".zattrs")) as fh:
            return tuple(json.load(fh).get("_ARRAY_DIMENSIONS", ()))
    except OSError:
        return ()
