"""Y4M -> device planes: the input side of the path (rav1e reads Y4M through src/bin/decoder/y4m.rs
into Frame<T> planes padded by LUMA_PADDING, src/frame/mod.rs:22-23; here the planes go to the device
once per frame with the same edge replication, b200_plane_upload = v_frame Plane::pad).

Only parsing and plumbing: no pixel arithmetic happens on the host.
"""
import numpy as np

LUMA_PADDING = 64 + 16 + 8          # src/frame/mod.rs:22-23

# colour space tag -> (bit depth, xdec, ydec, has chroma); the tags the y4m crate rav1e uses accepts
_CS = {"420": (8, 1, 1, True), "420jpeg": (8, 1, 1, True), "420mpeg2": (8, 1, 1, True), "420paldv": (8, 1, 1, True),
       "420p10": (10, 1, 1, True), "420p12": (12, 1, 1, True), "422": (8, 1, 0, True), "422p10": (10, 1, 0, True),
       "422p12": (12, 1, 0, True), "444": (8, 0, 0, True), "444p10": (10, 0, 0, True), "444p12": (12, 0, 0, True),
       "mono": (8, 0, 0, False), "mono12": (12, 0, 0, False)}


class Y4MError(ValueError):
    pass


def parse_header(line):
    """b'YUV4MPEG2 W64 H64 F25:1 Ip A16:9 C420jpeg ...' -> dict"""
    toks = line.strip().split()
    if not toks or toks[0] != b"YUV4MPEG2":
        raise Y4MError("not a YUV4MPEG2 stream")
    f = {"C": "420"}
    for t in toks[1:]:
        f[t[:1].decode()] = t[1:].decode()
    if "W" not in f or "H" not in f:
        raise Y4MError("missing W / H")
    cs = f["C"]
    if cs not in _CS:
        raise Y4MError(f"unsupported colour space {cs!r}")
    bd, xdec, ydec, chroma = _CS[cs]
    return {"width": int(f["W"]), "height": int(f["H"]), "bit_depth": bd, "xdec": xdec, "ydec": ydec,
            "chroma": chroma, "fps": f.get("F", "25:1"), "interlace": f.get("I", "p"), "raw": f}


def frames(path_or_bytes):
    """Yields (header, [Y, U, V]) per frame; planes are 2-D numpy arrays (uint8, or little-endian uint16
    for more than 8 bits) viewing the file's bytes."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    nl = data.index(b"\n")
    hdr = parse_header(data[:nl])
    w, h, bd = hdr["width"], hdr["height"], hdr["bit_depth"]
    bps = 1 if bd == 8 else 2
    cw, ch = (w + hdr["xdec"]) >> hdr["xdec"], (h + hdr["ydec"]) >> hdr["ydec"]
    sizes = [(h, w)] + ([(ch, cw)] * 2 if hdr["chroma"] else [])
    pos = nl + 1
    dt = np.uint8 if bps == 1 else np.dtype("<u2")
    while pos < len(data):
        end = data.index(b"\n", pos)
        if not data[pos:end].startswith(b"FRAME"):
            raise Y4MError(f"expected FRAME at byte {pos}")
        pos = end + 1
        planes = []
        for ph, pw in sizes:
            n = ph * pw * bps
            if pos + n > len(data):
                raise Y4MError("truncated frame")
            planes.append(np.frombuffer(data, dt, ph * pw, pos).reshape(ph, pw))
            pos += n
        yield hdr, planes


def upload_frame(ctx, planes, hdr, luma_pad=LUMA_PADDING):
    """[Y, U, V] -> device planes padded like Frame::new_with_padding (chroma padding = luma padding
    decimated); edges are replicated on the device."""
    out = []
    for i, p in enumerate(planes):
        pad = luma_pad if i == 0 else max(luma_pad >> hdr["xdec"], luma_pad >> hdr["ydec"])
        out.append(ctx.plane_from_host(np.ascontiguousarray(p), pad))
    return out
