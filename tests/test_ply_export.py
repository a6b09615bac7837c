"""HRBFFusion::savePly / GlobalModel::downloadMap: the oracle's byte-level restatement against a hand-computed
vertex (CPU), and the device-side filter + packing through the C ABI against the oracle (GPU, bit-exact)."""
import struct

import numpy as np
import pytest


def _surfel(pos, conf, rgb, submap, normal, radius, k1, k2):
    s = np.zeros(20, np.float32)
    s[0:3] = pos; s[3] = conf
    s[4] = float((rgb[0] << 16) + (rgb[1] << 8) + rgb[2]); s[5] = submap; s[6] = 1; s[7] = 2
    s[8:11] = normal; s[11] = radius
    s[15] = k1; s[19] = k2
    return s


def test_oracle_ply_known_answer(orc):
    a = _surfel((1.0, -2.0, 3.5), 7.0, (200, 100, 50), 3.0, (0.0, 0.6, -0.8), 0.0125, 4.5, -1.25)
    b = _surfel((9.0, 9.0, 9.0), 0.0, (1, 2, 3), 0.0, (1.0, 0.0, 0.0), 0.5, 0.0, 0.0)          # conf == threshold: NOT written (strict >)
    blob = orc.savePlyBytes(np.stack([a, b, a]), 0.0)
    head, body = blob.split(b"end_header\n")
    lines = head.decode().split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex 2"
    assert lines[3:6] == ["property float x", "property float y", "property float z"]
    assert lines[6:9] == ["property uchar red", "property uchar green", "property uchar blue"]
    assert lines[9:16] == ["property float nx", "property float ny", "property float nz", "property float curvature_max",
                           "property float curvature_min", "property float radius", "property float submapIndex"]
    assert len(body) == 2 * 43
    rec = struct.unpack("<3f3B7f", body[:43])
    want = (1.0, -2.0, 3.5, 200, 100, 50, -0.0, -0.6, 0.8, 4.5, -1.25, 0.0125, 3.0)
    assert rec == tuple(np.float32(x).item() if isinstance(x, float) else x for x in want)
    assert body[43:] == body[:43]
    assert orc.savePlyBytes(np.zeros((0, 20), np.float32)).split(b"end_header\n")[1] == b""


def test_ply_header_matches_oracle(orc):
    """hrbf_ply_header is host code: the library's header text against the oracle's, without a GPU"""
    import ctypes as C
    from hrbffusion3d_b200 import lib
    L = lib()
    L.hrbf_ply_header.restype = C.c_size_t
    for n in (0, 1, 305191, 4294967295):
        buf = C.create_string_buffer(512)
        k = L.hrbf_ply_header(C.c_uint(n), buf, C.c_size_t(512))
        want = orc.savePlyBytes(np.zeros((0, 20), np.float32)).replace(b"element vertex 0", b"element vertex %d" % n)
        assert buf.raw[:k] == want
    assert L.hrbf_ply_header(C.c_uint(5), C.create_string_buffer(64), C.c_size_t(64)) == 0      # too small: 0, nothing truncated silently
    assert L.hrbf_ply_header(C.c_uint(5), None, C.c_size_t(0)) == 0


@pytest.mark.gpu
def test_export_ply_and_download_map_match_oracle(orc, cuda, tmp_path):
    torch = cuda
    from hrbffusion3d_b200 import synth
    from hrbffusion3d_b200.fusion import GlobalModel, HRBFFusion
    from hrbffusion3d_b200._lib import HrbfError, check, lib
    import ctypes as C
    W, H = 320, 240
    cam = synth.default_camera(W, H)
    rng = np.random.default_rng(5)
    n = 70001                                                    # not a multiple of the 256-surfel tile
    s = rng.standard_normal((n, 20)).astype(np.float32)
    s[:, 3] = rng.uniform(0, 12, n).astype(np.float32)          # confidence
    s[:, 4] = rng.integers(0, 1 << 24, n).astype(np.float32)    # 24-bit colour in a float
    s[:, 5] = rng.integers(0, 40, n).astype(np.float32)
    s[::97, 3] = 5.0                                             # exactly on the threshold: dropped
    gm = GlobalModel(W, H, cam, capacity=1 << 17)
    gm.setModel(s)
    assert np.array_equal(gm.downloadMap(), s)
    for thr in (0.0, 5.0, 100.0):
        assert gm.exportPly(thr) == orc.savePlyBytes(s, thr), thr
    assert np.array_equal(gm.downloadMap(), s)                   # the export leaves the map alone
    # empty map, undersized buffers
    gm.setModel(s[:0])
    assert gm.exportPly(0.0) == orc.savePlyBytes(s[:0], 0.0) and gm.downloadMap().shape == (0, 20)
    gm.setModel(s[:1000])
    cnt = C.c_uint(0)
    small = np.zeros(43 * 10, np.uint8)
    with pytest.raises(HrbfError):
        check(lib().hrbf_model_export_ply(gm._h, C.c_float(0.0), small.ctypes.data_as(C.c_void_p), C.c_size_t(small.size), C.byref(cnt), None))
    with pytest.raises(HrbfError):
        check(lib().hrbf_model_download_map(gm._h, np.zeros(20, np.float32).ctypes.data_as(C.POINTER(C.c_float)), 1, C.byref(cnt), None))
    # after real frames: the file of the orchestrator, and the map keeps evolving normally afterwards
    sc = synth.Scene("room")
    fr = [synth.render_depth(sc, p, W, H, cam, noise=True, seed=i) for i, p in enumerate(synth.circle_trajectory(4, frames_per_rev=120))]
    a, b = HRBFFusion(W, H, cam, capacity=1 << 19), HRBFFusion(W, H, cam, capacity=1 << 19)
    for i, (depth, rgb) in enumerate(fr):
        a.processFrame(rgb, depth); b.processFrame(rgb, depth)
        if i == 2:
            path = tmp_path / "map.ply"
            a.savePly(str(path), 0.0)
            assert path.read_bytes() == orc.savePlyBytes(a.globalModel.downloadMap(), 0.0)
    assert np.array_equal(a.globalModel.downloadMap(), b.globalModel.downloadMap())      # savePly in between changed nothing
