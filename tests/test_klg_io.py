""".klg log reader / writer and the TUM trajectory writer (CPU)."""
import struct
import zlib

import numpy as np
import pytest

from hrbffusion3d_b200 import klg


def _frames(n, W, H, seed=0):
    rng = np.random.default_rng(seed)
    return [(1000 * i + 7, rng.integers(0, 6000, (H, W)).astype(np.uint16), rng.integers(0, 255, (H, W, 3)).astype(np.uint8)) for i in range(n)]


@pytest.mark.parametrize("compress", [True, False])
def test_klg_round_trip(compress):
    W, H = 64, 48
    fr = _frames(5, W, H)
    blob = klg.write_klg(fr, W, H, compress_depth=compress)
    assert struct.unpack_from("<i", blob, 0)[0] == 5
    rd = klg.KlgReader(blob, W, H)
    assert len(rd) == 5
    for (ts, d, c), (ts2, d2, c2) in zip(fr, rd):
        assert ts == ts2 and np.array_equal(d, d2) and np.array_equal(c, c2)
    # flipColors (RawLogReader.cpp:112-118)
    ts, d, c = klg.KlgReader(blob, W, H, flip_colors=True).next()
    assert np.array_equal(c, fr[0][2][..., ::-1])


def test_klg_layout_is_the_reference_wire_format():
    """hand-built log: int32 n | int64 ts, int32 depthSize, int32 imageSize, zlib depth, raw rgb | frame without image"""
    W, H = 8, 4
    depth = (np.arange(W * H, dtype=np.uint16) * 3).reshape(H, W)
    rgb = np.arange(W * H * 3, dtype=np.uint8).reshape(H, W, 3)
    z = zlib.compress(depth.tobytes())
    blob = struct.pack("<i", 2) + struct.pack("<qii", 123456789012, len(z), W * H * 3) + z + rgb.tobytes()
    blob += struct.pack("<qii", 5, W * H * 2, 0) + depth.tobytes()
    rd = klg.KlgReader(blob, W, H)
    ts, d, c = rd.next()
    assert ts == 123456789012 and np.array_equal(d, depth) and np.array_equal(c, rgb)
    ts, d, c = rd.next()
    assert ts == 5 and np.array_equal(d, depth) and not c.any()        # imageSize 0 -> black (RawLogReader.cpp:103-106)


def test_klg_errors_are_loud():
    W, H = 8, 4
    blob = klg.write_klg(_frames(2, W, H), W, H)
    with pytest.raises(ValueError):
        list(klg.KlgReader(blob[:-5], W, H))
    with pytest.raises(ValueError):
        klg.KlgReader(b"\x01", W, H)
    with pytest.raises(ValueError):
        klg.write_klg([(0, np.zeros((H, W + 1), np.uint16), np.zeros((H, W, 3), np.uint8))], W, H)
    assert len(klg.KlgReader(klg.write_klg([], W, H), W, H)) == 0


def test_tum_trajectory_format():
    c, s = np.cos(0.3), np.sin(0.3)
    P = np.eye(4)
    P[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    P[:3, 3] = [0.1, -0.2, 1.5]
    txt = klg.format_tum_trajectory([1305031102175304, 2000000], [P, np.concatenate([np.eye(3).ravel(), [1, 2, 3]])])
    l0, l1 = txt.strip().split("\n")
    f = l0.split()
    assert f[0] == "1305031102.175304" and len(f) == 8
    np.testing.assert_allclose([float(v) for v in f[1:4]], [0.1, -0.2, 1.5], rtol=1e-6)
    np.testing.assert_allclose([float(v) for v in f[4:]], [0, 0, np.sin(0.15), np.cos(0.15)], atol=1e-6)
    assert l1 == "2.000000 1 2 3 0 0 0 1"
    # quaternion branches: 180 degree rotations (trace <= 0)
    for axis in range(3):
        R = -np.eye(3); R[axis, axis] = 1
        q = klg.rotation_to_quaternion(R)
        e = [0, 0, 0, 0]; e[axis] = 1
        np.testing.assert_allclose(np.abs(q), e, atol=1e-12)


def test_klg_jpeg_image_branch_round_trip():
    """The JPEG branch of the .klg image payload (GUI/src/Tools/RawLogReader.cpp:103-105: imageSize != w*h*3 -> jpeg.readData): frames
    written with JPEG-compressed images read back as the decoder's RGB output (exactly what cv2.imdecode gives for the stored stream),
    depth stays lossless, a raw frame in the same log is still recognised by its size, flipColors swaps channels.  (libjpeg's headers
    are not installed, so this branch cannot be pinned to the reference's own reader like the raw / zlib branches are.)"""
    cv2 = pytest.importorskip("cv2")
    import struct
    W, H = 96, 64
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H, 0:W]
    frames = []
    for i in range(3):
        rgb = np.stack([(xx * 2 + 10 * i) % 256, (yy * 3) % 256, (xx + yy) % 256], -1).astype(np.uint8)
        depth = rng.integers(0, 20000, (H, W)).astype(np.uint16)
        frames.append((1000 * i, depth, rgb))
    blob = klg.write_klg(frames, W, H, jpeg_quality=90)
    assert len(blob) < sum(f[2].size for f in frames)                       # the images really are compressed
    out = list(klg.KlgReader(blob, W, H))
    assert len(out) == 3
    pos = 4
    for (ts, depth, rgb), (ts2, d2, r2) in zip(frames, out):
        assert ts2 == ts and np.array_equal(d2, depth)
        _, dsz, isz = struct.unpack_from("<qii", blob, pos)
        stream = np.frombuffer(blob, np.uint8, isz, pos + 16 + dsz)
        pos += 16 + dsz + isz
        assert isz != W * H * 3
        assert np.array_equal(r2, cv2.imdecode(stream, cv2.IMREAD_COLOR)[..., ::-1])      # the decoder's output, channel order RGB
        err = r2.astype(np.float64) - rgb
        assert 10 * np.log10(255.0 ** 2 / np.mean(err ** 2)) > 30.0                      # and close to what was stored (lossy codec)
    flipped = list(klg.KlgReader(blob, W, H, flip_colors=True))
    assert np.array_equal(flipped[0][2], out[0][2][..., ::-1])
    # a truncated JPEG stream is an error, not a black frame
    with pytest.raises(ValueError):
        bad = bytearray(klg.write_klg(frames[:1], W, H, jpeg_quality=90))
        _, dsz, isz = struct.unpack_from("<qii", bad, 4)
        bad[4 + 16 + dsz:4 + 16 + dsz + isz] = bytes(isz)
        list(klg.KlgReader(bytes(bad), W, H))
