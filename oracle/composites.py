"""Oracle ratio sharpening (numpy) -- TEST INFRASTRUCTURE ONLY.

Restates ``RatioSharpenedRGB`` / ``SelfSharpenedRGB`` (satpy/composites/resolution.py:32-233: ``_get_sharpening_ratio``
:157-165, ``_mean4`` :168-192, the band products :143-154) and the common channel mask of ``GenericCompositor``
(satpy/composites/core.py:489-490) for a single chunk.  Pinned by the reference's golden vectors
(satpy/tests/compositor_tests/test_resolution.py:143-214) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

COLORS = ["red", "green", "blue"]


def sharpening_ratio(high_res, low_res):
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = high_res / low_res
    ratio[~np.isfinite(ratio) | (ratio < 0)] = 1.0
    np.clip(ratio, 0, 1.5, out=ratio)
    return ratio


def mean4(data, offset=(0, 0)):
    rows, cols = data.shape
    row_offset, col_offset = offset[0] % 2, offset[1] % 2
    row_after = (row_offset + rows) % 2
    col_after = (col_offset + cols) % 2
    pad = ((row_offset, row_after), (col_offset, col_after))
    rows2 = rows + row_offset + row_after
    cols2 = cols + col_offset + col_after
    av_data = np.pad(data, pad, "edge")
    new_shape = (int(rows2 / 2.), 2, int(cols2 / 2.), 2)
    with np.errstate(invalid="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            data_mean = np.nanmean(av_data.reshape(new_shape), axis=(1, 3))
    data_mean = np.repeat(np.repeat(data_mean, 2, axis=0), 2, axis=1)
    return data_mean[row_offset:row_offset + rows, col_offset:col_offset + cols]


def ratio_sharpened_rgb(red, green, blue, high=None, high_resolution_band="red", neutral_resolution_band=None,
                        self_sharpen=False, offset=(0, 0), common_channel_mask=True):
    bands = {"red": np.array(red), "green": np.array(green), "blue": np.array(blue)}
    if self_sharpen:
        high = bands[high_resolution_band]
        bands[high_resolution_band] = mean4(high, offset)
    if high is not None and high_resolution_band is not None:
        ratio = sharpening_ratio(np.array(high), bands[high_resolution_band])
        bands[high_resolution_band] = np.array(high)
        for color in COLORS:
            if color != neutral_resolution_band and color != high_resolution_band:
                bands[color] = bands[color] * ratio
    data = np.stack([bands[c] for c in COLORS])
    if common_channel_mask:
        data = np.where(~np.isnan(data).any(axis=0), data, np.nan)
    return data
