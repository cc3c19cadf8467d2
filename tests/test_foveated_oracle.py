"""CPU: the oracle's foveated stereo copy (oracle/gsm_oracle_copy.c). The reference holds no vector
for this copy (parity unpinned, see the file header); what is pinned here is that at 1:1 it IS the literal row-flipped copy
the sideBySide path is checked against, plus hand-computed known answers for the resampling and the format conversions."""
import numpy as np

from tests import foveation_util as fv


def _halfs(a):
    return np.asarray(a, np.float16).view(np.uint16)


def test_one_to_one_is_the_literal_copy(oracle):
    rng = np.random.default_rng(3)
    for H, W in ((37, 53), (270, 481)):
        c2 = (rng.standard_normal((2, H, W, 4)) * np.exp(rng.uniform(-8, 8, (2, H, W, 4)))).astype(np.float16).view(np.uint16)
        c2[0, 3, 4, :] = [0x7C00, 0xFC00, 0x7E00, 0x8000]  # +inf, -inf, NaN, -0 must survive a 1:1 copy
        for flip in (0, 1):
            out = oracle.stereo_copy_foveated(c2, flip, 2 * W, H, 1, 0, ((0, 0, W, H), (W, 0, W, H)))
            got = out.reshape(H, 2 * W * 8).view(np.uint16).reshape(H, 2 * W, 4)
            ref = np.empty((H, 2 * W, 4), np.uint16)
            oracle.lib().gsmo_stereo_copy(oracle._p(c2), W, H, flip, oracle._p(ref))
            assert np.array_equal(got, ref)


def test_layered_one_to_one_and_untouched_texels(oracle):
    rng = np.random.default_rng(4)
    H, W = 20, 24
    c2 = rng.uniform(0, 1, (2, H, W, 4)).astype(np.float16).view(np.uint16)
    # drawable larger than the viewports: a 4-texel border stays as it was
    out = oracle.stereo_copy_foveated(c2, 0, W + 8, H + 8, 2, 0, ((4, 4, W, H), (4, 4, W, H)))
    img = out.reshape(2, H + 8, (W + 8) * 8).view(np.uint16).reshape(2, H + 8, W + 8, 4)
    assert np.array_equal(img[:, 4:4 + H, 4:4 + W], c2)
    border = np.ones((H + 8, W + 8), bool)
    border[4:4 + H, 4:4 + W] = False
    assert (out.reshape(2, H + 8, W + 8, 8)[:, border] == 0xAB).all()


def test_bilinear_known_answers(oracle):
    # 2x1 source per eye, magnified 4x horizontally: texel centres at u = 0.25 and 0.75
    src = np.zeros((2, 1, 2, 4), np.float16)
    src[0, 0, 0, :] = [0.0, 1.0, 2.0, 1.0]
    src[0, 0, 1, :] = [1.0, 3.0, -2.0, 0.0]
    src[1] = src[0] * 2
    out = oracle.stereo_copy_foveated(src.view(np.uint16), 0, 8, 1, 2, 0, ((0, 0, 8, 1), (0, 0, 8, 1)))
    img = out.reshape(2, 1, 8 * 8).view(np.float16).reshape(2, 1, 8, 4).astype(np.float32)
    # dest centres x + 0.5 -> u = (x + 0.5) / 8 -> tx = 2u - 0.5: -0.375 (clamped: texel 0), -0.125, 0.125, 0.375, 0.625, 0.875, 1.125 ...
    w = np.array([0, 0, 0.125, 0.375, 0.625, 0.875, 1, 1], np.float32)
    for k in range(4):
        a, b = float(src[0, 0, 0, k]), float(src[0, 0, 1, k])
        want = (a + w * (b - a)).astype(np.float16).astype(np.float32)
        assert np.array_equal(img[0, 0, :, k], want)
        assert np.array_equal(img[1, 0, :, k], (2 * (a + w * (b - a))).astype(np.float16).astype(np.float32))


def test_flip_and_minification(oracle):
    # 1x4 column minified to 2 rows: each output row is the mean of two source rows; flipY reverses them
    src = np.zeros((2, 4, 1, 4), np.float16)
    src[:, :, 0, 0] = [1, 3, 5, 7]
    for flip, want in ((0, [2, 6]), (1, [6, 2])):
        out = oracle.stereo_copy_foveated(src.view(np.uint16), flip, 1, 2, 2, 0, ((0, 0, 1, 2), (0, 0, 1, 2)))
        img = out.reshape(2, 2, 8).view(np.float16).reshape(2, 2, 1, 4)
        assert img[0, :, 0, 0].astype(np.float32).tolist() == want


def test_attachment_formats(oracle):
    vals = np.array([[0.0, 0.5, 1.0, 0.25], [-1.0, 2.0, np.nan, 0.5], [0.0031308, 0.2158, 0.7305, 1.0], [1 / 255, 0.5 / 255, 0.1, 0.9]],
                    np.float16)
    src = np.zeros((2, 1, 4, 4), np.float16)
    src[0, 0] = vals
    src[1, 0] = vals
    vp = ((0, 0, 4, 1), (0, 0, 4, 1))

    def srgb(c):
        c = float(np.float16(c))
        if not c > 0:
            return 0
        if c >= 1:
            return 255
        s = 12.92 * c if c <= 0.0031308 else 1.055 * c ** (1 / 2.4) - 0.055
        return int(np.floor(s * 255 + 0.5))

    def un(c):
        c = float(np.float16(c))
        if not c > 0:
            return 0
        if c >= 1:
            return 255
        return int(np.rint(np.float32(c) * np.float32(255.0)))

    for fmt, order, enc in ((1, (2, 1, 0), un), (2, (2, 1, 0), srgb), (3, (0, 1, 2), un), (4, (0, 1, 2), srgb)):
        out = oracle.stereo_copy_foveated(src.view(np.uint16), 0, 4, 1, 2, fmt, vp).reshape(2, 1, 4, 4)
        for x in range(4):
            want = [enc(vals[x, order[0]]), enc(vals[x, order[1]]), enc(vals[x, order[2]]), un(vals[x, 3])]
            assert out[0, 0, x].tolist() == want, (fmt, x)
    # the classic sRGB anchors
    assert srgb(0.5) == 188 and srgb(0.2158) == 128 and srgb(1.0) == 255 and srgb(0.0) == 0


def test_rate_map_reads_the_screen_position_of_each_physical_texel(oracle):
    # a horizontal ramp: value == screen x; through a rate map every physical texel must hold (about) its screen x
    W, H = 64, 8
    ramp = np.tile((np.arange(W, dtype=np.float32) + 0.5)[None, :, None], (H, 1, 4)).astype(np.float16)
    c2 = np.stack([ramp, ramp]).view(np.uint16)
    sx, sy = fv.layer(W, H, fv.FOVEATED_H, (1.0,))
    out = oracle.stereo_copy_foveated(c2, 0, sx.size, sy.size, 2, 0, ((0, 0, W, H), (0, 0, W, H)), rate_layers=[(sx, sy)])
    img = out.reshape(2, sy.size, sx.size * 8).view(np.float16).reshape(2, sy.size, sx.size, 4).astype(np.float32)
    inner = (sx >= 0.5) & (sx <= W - 0.5)
    assert np.abs(img[0, 0, inner, 0] - sx[inner]).max() <= 0.03  # 8-bit sub-texel weights + half rounding
    assert sx.size < W  # the physical image is smaller than the screen
