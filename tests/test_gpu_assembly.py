"""Assembling a row-tiled image from column spans (satpy_b200/parallel.py) -- ``-m gpu``.

The multi-rank exchange is emulated on one GPU: every "rank" owns an image, searches its blocks, copies only the
reachable column span of each block into the other image (``b200_copy_2d``) and writes the fill value outside the
spans of the blocks it does not own (``b200_fill_f32_rects``).  Both images must equal the single-pass result bit for
bit -- which also proves that ``block_column_span`` never cuts a hit off at full resolution geometry ratios."""
import ctypes as C

import numpy as np
import pytest

from satpy_b200 import _lib, demo, parallel
from satpy_b200.resample import B200NearestResampler

pytestmark = pytest.mark.gpu


def test_fill_rects_and_copy_2d_match_numpy():
    import torch
    lib = _lib.load()
    h, w = 301, 1003                 # odd width: rows start at every alignment
    img = torch.full((h, w), 5.0, dtype=torch.float32, device="cuda")
    ref = np.full((h, w), 5.0, np.float32)
    rects = [(0, 10, 0, 17), (10, 100, 900, 103), (150, 1, 3, 999), (200, 101, 0, w), (120, 5, 500, 0)]
    flat = (C.c_int64 * (4 * len(rects)))(*[v for r in rects for v in r])
    _lib.check(lib.b200_fill_f32_rects(img.data_ptr(), w, flat, len(rects), float("nan"), _lib.stream_ptr(torch)))
    for r0, n, c0, nc in rects:
        ref[r0:r0 + n, c0:c0 + nc] = np.nan
    np.testing.assert_array_equal(img.cpu().numpy(), ref)
    _lib.check(lib.b200_fill_f32_2d(img[20, 7:].data_ptr(), w, 33, 4, -1.0, _lib.stream_ptr(torch)))
    ref[20:24, 7:40] = -1.0
    np.testing.assert_array_equal(img.cpu().numpy(), ref)
    src = torch.arange(h * w, dtype=torch.float32, device="cuda").reshape(h, w)
    _lib.check(lib.b200_copy_2d(img[50, 100:].data_ptr(), w * 4, src[50, 100:].data_ptr(), w * 4, 260 * 4, 30,
                                _lib.stream_ptr(torch)))
    ref[50:80, 100:360] = src.cpu().numpy()[50:80, 100:360]
    np.testing.assert_array_equal(img.cpu().numpy(), ref)
    bad = (C.c_int64 * 4)(0, 1, w - 2, 5)
    assert lib.b200_fill_f32_rects(img.data_ptr(), w, bad, 1, 0.0, _lib.stream_ptr(torch)) == _lib.B200_EINVAL


@pytest.mark.parametrize("world,nsub", [(2, 3), (4, 2)])
def test_span_exchange_reassembles_the_image_on_every_rank(world, nsub):
    import torch
    lib = _lib.load()
    s, d = 900 / 10848, 3204 / 80150            # 12 km source pixels, 6.2 km target pixels: the config-5 ratio
    src, dst = demo.make_area("goes_east_abi_f_1km", s), demo.make_area("global_eqc_500m", d)
    radius = 2.0 * 1002.0 / s
    data = torch.rand(src.shape, device="cuda", dtype=torch.float32) * 100
    full = B200NearestResampler(src, dst).resample(data, radius_of_influence=radius)
    block_rows, blocks = parallel.row_tiles(dst.height, world * nsub)
    spans = [parallel.block_column_span(src, dst, r0, n, radius) for r0, n in blocks]
    assert sum((c1 - c0) * n for (c0, c1), (_, n) in zip(spans, blocks)) < 0.5 * dst.size
    images = [torch.full((world * nsub * block_rows, dst.width), -7.0, dtype=torch.float32, device="cuda")
              for _ in range(world)]
    stream = _lib.stream_ptr(torch)
    for rank in range(world):
        mine = images[rank]
        rects = []
        for b, ((r0, n), (c0, c1)) in enumerate(zip(blocks, spans)):
            if b % world == rank or n == 0:
                continue
            slot = b * block_rows
            rects += [slot, n, 0, c0] if c1 > c0 else [slot, n, 0, dst.width]
            if c1 > c0 and c1 < dst.width:
                rects += [slot, n, c1, dst.width - c1]
        flat = (C.c_int64 * len(rects))(*rects)
        _lib.check(lib.b200_fill_f32_rects(mine.data_ptr(), dst.width, flat, len(rects) // 4, float("nan"), stream))
        for b, ((r0, n), (c0, c1)) in enumerate(zip(blocks, spans)):
            if b % world != rank or n == 0:
                continue
            slot = b * block_rows
            s0, sn = parallel.source_row_window(src, dst, r0, n, radius)
            if sn == 0:
                mine[slot:slot + n].fill_(float("nan"))
            else:
                sub = src.row_window(s0, sn)
                B200NearestResampler(sub, dst).resample(data[s0:s0 + sn].contiguous(), radius_of_influence=radius,
                                                        row_window=(r0, n), out=mine[slot:slot + n])
            for peer in range(world):
                if peer == rank or c1 <= c0:
                    continue
                _lib.check(lib.b200_copy_2d(images[peer][slot, c0:].data_ptr(), dst.width * 4,
                                            mine[slot, c0:].data_ptr(), dst.width * 4, (c1 - c0) * 4, n, stream))
    ref = full.cpu().numpy()
    for rank in range(world):
        got = images[rank].cpu().numpy()
        for b, (r0, n) in enumerate(blocks):
            np.testing.assert_array_equal(got[b * block_rows: b * block_rows + n], ref[r0:r0 + n],
                                          err_msg=f"rank {rank}, block {b}")


@pytest.mark.parametrize("n", [5, 4096, 3 * 4096 + 1234, 1 << 20])
def test_calibration_and_gather_variants_are_bit_identical(n):
    """The lane-interleaved and the bulk-asynchronous (cp.async.bulk + mbarrier) kernels against the first release's,
    on sizes with ragged tails (less than one tile, whole tiles + a tail)."""
    import torch
    lib = _lib.load()
    stream = _lib.stream_ptr(torch)
    rng = np.random.default_rng(n)
    counts = torch.from_numpy(rng.integers(0, 4096, size=n).astype(np.int16)).cuda()
    for band, calib in (("C02", "radiance"), ("C02", "reflectance"), ("C13", "brightness_temperature")):
        st = demo.abi_calibration(band).struct(calib)
        outs = []
        for variant in (0, 1, 2):
            out = torch.full((n,), -3.0, dtype=torch.float32, device="cuda")
            _lib.check(lib.b200_abi_calibrate_variant(counts.data_ptr(), out.data_ptr(), n, C.byref(st), variant, stream))
            outs.append(out.cpu().numpy())
        np.testing.assert_array_equal(outs[1], outs[0])
        np.testing.assert_array_equal(outs[2], outs[0])
    n_src = 50000
    data = torch.from_numpy(rng.uniform(0, 100, size=n_src).astype(np.float32)).cuda()
    gi = torch.from_numpy(rng.integers(-1, n_src, size=n).astype(np.int32)).cuda()
    fill = (C.c_char * 4).from_buffer_copy(np.float32(np.nan).tobytes())
    res = []
    for variant in (0, 1):
        out = torch.full((n,), -3.0, dtype=torch.float32, device="cuda")
        _lib.check(lib.b200_nn_gather_variant(data.data_ptr(), out.data_ptr(), gi.data_ptr(), n, 1, n_src, n, 4,
                                              C.cast(fill, C.c_void_p), variant, stream))
        res.append(out.cpu().numpy())
    np.testing.assert_array_equal(res[1], res[0])
    exp = np.where(gi.cpu().numpy() < 0, np.nan, data.cpu().numpy()[np.maximum(gi.cpu().numpy(), 0)])
    np.testing.assert_array_equal(res[0], exp.astype(np.float32))
    # three bands through one index stream (band strides that keep 16-byte alignment, and one that does not)
    data3 = torch.from_numpy(rng.uniform(0, 100, size=(3, n_src)).astype(np.float32)).cuda()
    for pad in (0, 4, 2):
        stride = n + pad
        res3 = []
        for variant in (0, 1):
            out = torch.full((3 * stride,), -3.0, dtype=torch.float32, device="cuda")
            _lib.check(lib.b200_nn_gather_variant(data3.data_ptr(), out.data_ptr(), gi.data_ptr(), n, 3, n_src, stride, 4,
                                                  C.cast(fill, C.c_void_p), variant, stream))
            res3.append(out.cpu().numpy())
        np.testing.assert_array_equal(res3[1], res3[0])
        for b in range(3):
            e = np.where(gi.cpu().numpy() < 0, np.nan, data3[b].cpu().numpy()[np.maximum(gi.cpu().numpy(), 0)])
            np.testing.assert_array_equal(res3[0][b * stride: b * stride + n], e.astype(np.float32))
