"""Dependency-free Float32 GeoTIFF writer for the output side of the match path.

The reference writes rasters through GDAL (``BaseSpatialGrid.save``, dem.py:291-306: one
Float32 band, geotransform, projection) and publishes search results as 4-band rasters
``1 = amplitude, 2 = relative age, 3 = orientation, 4 = SNR`` (CHANGELOG.md:20).  GDAL is
not a dependency here: this module writes the same files with ``struct`` and NumPy --
baseline TIFF 6.0 (or BigTIFF beyond 4 GB), uncompressed strips, ``SampleFormat`` IEEE float,
pixel-interleaved bands (GDAL's GTiff default), and the GeoTIFF tags ``ModelPixelScale`` /
``ModelTiepoint`` / ``GeoKeyDirectory`` derived from a GDAL-style geotransform.
"""
import struct

import numpy as np

BAND_ORDER = ("amplitude", "age", "orientation", "snr")      # CHANGELOG.md:20; = the [amp, age, angle, snr] stack

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}


def _geo_tags(geo_transform, epsg, projected):
    """ModelPixelScale (33550), ModelTiepoint (33922), GeoKeyDirectory (34735) for a north-up
    geotransform ``(x0, dx, 0, y0, 0, dy)`` as ``gdal.GetGeoTransform`` returns it
    (dem.py:323-329); dy is negative for north-up rasters."""
    x0, dx, rx, y0, ry, dy = (float(v) for v in geo_transform)
    if rx != 0.0 or ry != 0.0:
        raise ValueError("rotated geotransforms are not supported")
    tags = [(33550, 12, (abs(dx), abs(dy), 0.0)),
            (33922, 12, (0.0, 0.0, 0.0, x0, y0, 0.0))]
    keys = [(1024, 0, 1, 1 if projected else 2),         # GTModelTypeGeoKey: projected / geographic
            (1025, 0, 1, 1)]                             # GTRasterTypeGeoKey: PixelIsArea
    if epsg is not None:
        keys.append((3072 if projected else 2048, 0, 1, int(epsg)))   # ProjectedCSType / GeographicType
    flat = [1, 1, 0, len(keys)]
    for k in sorted(keys):
        flat.extend(k)
    tags.append((34735, 3, tuple(flat)))
    return tags


def write_geotiff(filename, bands, geo_transform=None, epsg=None, projected=True, nodata=None,
                  rows_per_strip=None):
    """Write ``bands`` -- (ny, nx) or (nbands, ny, nx), any real dtype -- as a Float32 GeoTIFF.

    ``geo_transform``: GDAL-style 6-tuple (``GeorefInfo.geo_transform``, dem.py:323); without
    it the file is a plain TIFF, which GDAL opens with dx = dy = 1.  ``nodata`` is stored in
    the ``GDAL_NODATA`` tag (42113).  Returns the number of bytes written."""
    a = np.asarray(bands)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("bands must be (ny, nx) or (nbands, ny, nx)")
    nb, ny, nx = a.shape
    row_bytes = nx * nb * 4
    if rows_per_strip is None:
        rows_per_strip = max(1, min(ny, (8 << 20) // max(row_bytes, 1)))
    n_strips = (ny + rows_per_strip - 1) // rows_per_strip
    data_bytes = ny * row_bytes
    big = data_bytes + (1 << 20) >= (1 << 32)

    tags = [(256, 4, (nx,)), (257, 4, (ny,)),
            (258, 3, (32,) * nb),                        # BitsPerSample
            (259, 3, (1,)),                              # Compression: none
            (262, 3, (1,)),                              # Photometric: BlackIsZero
            (277, 3, (nb,)),                             # SamplesPerPixel
            (278, 4, (rows_per_strip,)),
            (284, 3, (1,)),                              # PlanarConfiguration: chunky (pixel-interleaved)
            (339, 3, (3,) * nb)]                         # SampleFormat: IEEE floating point
    if nb > 1:
        tags.append((338, 3, (0,) * (nb - 1)))           # ExtraSamples: unspecified
    if geo_transform is not None:
        tags.extend(_geo_tags(geo_transform, epsg, projected))
    if nodata is not None:
        txt = (repr(float(nodata)) + "\0").encode("ascii")
        tags.append((42113, 2, tuple(bytes([c]) for c in txt)))
    off_type = 16 if big else 4
    tags.append((273, off_type, (0,) * n_strips))       # StripOffsets (patched below)
    counts = [min(rows_per_strip, ny - s * rows_per_strip) * row_bytes for s in range(n_strips)]
    tags.append((279, off_type, tuple(counts)))          # StripByteCounts
    tags.sort()

    # layout: header | IFD | out-of-line tag values | pixel data
    hdr = 16 if big else 8
    entry = 20 if big else 12
    inline = 8 if big else 4
    ifd_size = (8 if big else 2) + entry * len(tags) + (8 if big else 4)
    payloads = []
    cursor = hdr + ifd_size
    for tag, typ, vals in tags:
        fmt = "<%d%s" % (len(vals), _TYPES[typ])
        raw = struct.pack(fmt, *vals)
        if len(raw) <= inline:
            payloads.append((raw, None))
        else:
            cursor += cursor & 1
            payloads.append((raw, cursor))
            cursor += len(raw)
    data_start = (cursor + 15) // 16 * 16
    offsets = []
    o = data_start
    for c in counts:
        offsets.append(o)
        o += c
    with open(filename, "wb") as f:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, hdr) if big else struct.pack("<2sHI", b"II", 42, hdr))
        f.write(struct.pack("<Q" if big else "<H", len(tags)))
        blobs = []
        for (tag, typ, vals), (raw, where) in zip(tags, payloads):
            if tag == 273:
                raw = struct.pack("<%d%s" % (n_strips, _TYPES[typ]), *offsets)
            if where is None:
                field = raw.ljust(inline, b"\0")
            else:
                field = struct.pack("<Q" if big else "<I", where)
                blobs.append((where, raw))
            f.write(struct.pack("<HHQ" if big else "<HHI", tag, typ, len(vals)) + field)
        f.write(struct.pack("<Q" if big else "<I", 0))              # no further IFD
        for where, raw in blobs:
            f.seek(where)
            f.write(raw)
        f.seek(data_start)
        for s in range(n_strips):
            r0 = s * rows_per_strip
            r1 = min(ny, r0 + rows_per_strip)
            block = np.ascontiguousarray(np.moveaxis(a[:, r0:r1, :], 0, -1), dtype="<f4")
            f.write(block.tobytes())
        return f.tell()


def write_results(filename, results, georef=None, epsg=None, projected=True):
    """Search results as the reference publishes them (CHANGELOG.md:20): a 4-band Float32
    GeoTIFF, band order amplitude, age, orientation, SNR -- the order of the
    ``[amp, age, angle, snr]`` stack ``match`` returns (core.py:190-193).  ``georef``: a
    ``GeorefInfo`` (``.geo_transform``, or ``.dx/.dy`` and an upper-left corner)."""
    stack = np.asarray(results)
    if stack.ndim != 3 or stack.shape[0] != 4:
        raise ValueError("results must be the (4, ny, nx) stack [amp, age, angle, snr]")
    return write_geotiff(filename, stack, geo_transform=_transform_of(georef), epsg=epsg, projected=projected)


def _transform_of(georef):
    if georef is None:
        return None
    gt = getattr(georef, "geo_transform", None)
    if gt is not None:
        return tuple(gt)
    dx, dy = getattr(georef, "dx", None), getattr(georef, "dy", None)
    if dx is None:
        return None
    ulx = getattr(georef, "ulx", None)
    uly = getattr(georef, "uly", None)
    return (0.0 if ulx is None else ulx, dx, 0.0, 0.0 if uly is None else uly, 0.0, -abs(dy if dy is not None else dx))


def read_geotiff(filename):
    """Minimal reader for files written by ``write_geotiff`` (and any uncompressed, stripped,
    chunky Float32 TIFF): returns ``(bands[nb, ny, nx] float32, tags{id: tuple})``.  Used by
    the tests and by ``load_results``."""
    with open(filename, "rb") as f:
        buf = f.read()
    if buf[:2] != b"II":
        raise ValueError("little-endian TIFF expected")
    magic = struct.unpack_from("<H", buf, 2)[0]
    big = magic == 43
    if big:
        ifd = struct.unpack_from("<Q", buf, 8)[0]
        n = struct.unpack_from("<Q", buf, ifd)[0]
        pos, entry, inline = ifd + 8, 20, 8
    else:
        ifd = struct.unpack_from("<I", buf, 4)[0]
        n = struct.unpack_from("<H", buf, ifd)[0]
        pos, entry, inline = ifd + 2, 12, 4
    size = {1: 1, 2: 1, 3: 2, 4: 4, 12: 8, 16: 8}
    tags = {}
    for i in range(n):
        base = pos + i * entry
        tag, typ, cnt = struct.unpack_from("<HHQ" if big else "<HHI", buf, base)
        nbytes = size[typ] * cnt
        field = base + (12 if big else 8)
        where = field if nbytes <= inline else struct.unpack_from("<Q" if big else "<I", buf, field)[0]
        tags[tag] = struct.unpack_from("<%d%s" % (cnt, _TYPES[typ]), buf, where)
    nx, ny, nb = tags[256][0], tags[257][0], tags.get(277, (1,))[0]
    if tags.get(259, (1,))[0] != 1 or tags.get(284, (1,))[0] != 1 or set(tags[339]) != {3} or set(tags[258]) != {32}:
        raise ValueError("only uncompressed chunky Float32 TIFFs are supported")
    out = np.empty((ny, nx, nb), dtype="<f4")
    rps = tags[278][0]
    for s, (o, c) in enumerate(zip(tags[273], tags[279])):
        r0 = s * rps
        rows = c // (nx * nb * 4)
        out[r0:r0 + rows] = np.frombuffer(buf, dtype="<f4", count=rows * nx * nb, offset=o).reshape(rows, nx, nb)
    return np.moveaxis(out, -1, 0), tags
