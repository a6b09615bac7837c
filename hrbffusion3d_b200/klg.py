"""Input / output formats either side of the hot path.

.klg log  (GUI/src/Tools/RawLogReader.cpp:28-118):
    int32 numFrames; then per frame  int64 timestamp, int32 depthSize, int32 imageSize, depth bytes, image bytes
    depth: raw uint16[h][w] when depthSize == w*h*2, otherwise a zlib stream of it
    image: raw RGB8[h][w][3] when imageSize == w*h*3, JPEG otherwise (decoded with OpenCV when it is installed), absent
           (imageSize == 0) -> black
TUM trajectory  (Core/src/Utils/TrajectoryManager.cpp:313-343):  "timestamp tx ty tz qx qy qz qw", timestamp = us / 1e6 with 6 decimals
"""
import io
import struct
import zlib

import numpy as np


def write_klg(frames, width, height, compress_depth=True, jpeg_quality=None):
    """frames: iterable of (timestamp_us, depth uint16 [h,w], rgb uint8 [h,w,3]).  Returns the log as bytes.
    jpeg_quality = 1..100: the image is stored as a JPEG stream (what the reference's Logger writes for live captures and its
    RawLogReader hands to libjpeg, RawLogReader.cpp:103-105 / JPEGLoader.h); None: raw RGB8."""
    body = io.BytesIO()
    n = 0
    for ts, depth, rgb in frames:
        depth = np.ascontiguousarray(depth, np.uint16)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        if depth.shape != (height, width) or rgb.shape != (height, width, 3):
            raise ValueError("frame %d: shape mismatch" % n)
        d = depth.tobytes()
        if compress_depth:
            z = zlib.compress(d, 1)
            if len(z) != len(d):                # a stream of exactly w*h*2 bytes would be read back as raw
                d = z
        img = rgb.tobytes()
        if jpeg_quality is not None:
            import cv2
            ok, enc = cv2.imencode(".jpg", rgb[..., ::-1], [int(cv2.IMWRITE_JPEG_QUALITY), int(jpeg_quality)])
            if not ok:
                raise RuntimeError("frame %d: JPEG encoding failed" % n)
            if enc.size != rgb.size:            # a stream of exactly w*h*3 bytes would be read back as raw
                img = enc.tobytes()
        body.write(struct.pack("<qii", int(ts), len(d), len(img)))
        body.write(d)
        body.write(img)
        n += 1
    return struct.pack("<i", n) + body.getvalue()


class KlgReader:
    """Sequential reader over a .klg byte buffer (bytes / memoryview / file path)."""

    def __init__(self, source, width, height, flip_colors=False):
        if isinstance(source, str):
            with open(source, "rb") as f:
                source = f.read()
        self.buf = memoryview(source)
        self.width, self.height, self.flip = width, height, flip_colors
        if len(self.buf) < 4:
            raise ValueError("klg: truncated header")
        (self.num_frames,) = struct.unpack_from("<i", self.buf, 0)
        if self.num_frames < 0:
            raise ValueError("klg: negative frame count")
        self.pos, self.index = 4, 0

    def __len__(self):
        return self.num_frames

    def __iter__(self):
        while self.index < self.num_frames:
            yield self.next()

    def next(self):
        P = self.width * self.height
        if self.pos + 16 > len(self.buf):
            raise ValueError("klg: truncated frame header at frame %d" % self.index)
        ts, dsz, isz = struct.unpack_from("<qii", self.buf, self.pos)
        self.pos += 16
        if dsz < 0 or isz < 0 or self.pos + dsz + isz > len(self.buf):
            raise ValueError("klg: truncated frame %d" % self.index)
        draw = self.buf[self.pos:self.pos + dsz]
        iraw = self.buf[self.pos + dsz:self.pos + dsz + isz]
        self.pos += dsz + isz
        if dsz == P * 2:
            depth = np.frombuffer(draw, np.uint16).reshape(self.height, self.width).copy()
        else:
            d = zlib.decompress(bytes(draw))
            if len(d) != P * 2:
                raise ValueError("klg: depth of frame %d inflates to %d bytes, expected %d" % (self.index, len(d), P * 2))
            depth = np.frombuffer(d, np.uint16).reshape(self.height, self.width).copy()
        if isz == P * 3:
            rgb = np.frombuffer(iraw, np.uint8).reshape(self.height, self.width, 3).copy()
        elif isz > 0:
            try:
                import cv2
            except ImportError as e:           # the reference links OpenCV for this (JPEGLoader.h)
                raise RuntimeError("klg: JPEG-compressed image but OpenCV is not installed") from e
            bgr = cv2.imdecode(np.frombuffer(iraw, np.uint8), cv2.IMREAD_COLOR)
            if bgr is None or bgr.shape != (self.height, self.width, 3):
                raise ValueError("klg: cannot decode the image of frame %d" % self.index)
            rgb = bgr[..., ::-1].copy()
        else:
            rgb = np.zeros((self.height, self.width, 3), np.uint8)
        if self.flip:
            rgb = rgb[..., ::-1].copy()
        self.index += 1
        return ts, depth, rgb


def rotation_to_quaternion(R):
    """(x, y, z, w) of a 3x3 rotation, w >= 0 branch selection as Eigen::Quaternionf(Matrix3f) does"""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        x, y, z = (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q = [0.0, 0.0, 0.0]
        q[i] = 0.5 * s
        s = 0.5 / s
        w = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
        x, y, z = q
    return float(x), float(y), float(z), float(w)


def format_tum_trajectory(timestamps_us, poses):
    """poses: iterable of 4x4 (or [12] = R row-major, t) camera-to-world matrices.  Returns the file contents."""
    lines = []
    for ts, P in zip(timestamps_us, poses):
        P = np.asarray(P, np.float64)
        if P.size == 12:
            R, t = P[:9].reshape(3, 3), P[9:12]
        else:
            P = P.reshape(4, 4)
            R, t = P[:3, :3], P[:3, 3]
        x, y, z, w = rotation_to_quaternion(R)
        # the reference streams floats with operator<< (6 significant digits), the timestamp fixed with 6 decimals
        lines.append("%.6f %s %s %s %s %s %s %s" % (float(ts) / 1e6, *("%g" % v for v in (np.float32(t[0]), np.float32(t[1]), np.float32(t[2]), np.float32(x), np.float32(y), np.float32(z), np.float32(w)))))
    return "\n".join(lines) + ("\n" if lines else "")


def write_tum_trajectory(path, timestamps_us, poses):
    with open(path, "w") as f:
        f.write(format_tum_trajectory(timestamps_us, poses))
