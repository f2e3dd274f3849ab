"""BGZF writer used by the tests and the benchmark of the device inflater (SAM specification section 4.1): payload blocks of
at most 0xff00 bytes, each a gzip member with the 'BC' extra subfield, and the 28-byte end-of-file marker."""
from __future__ import annotations

import struct
import zlib

EOF_MARKER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def compress(data: bytes, level: int = 6, block: int = 0xff00, eof: bool = True, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    out = []
    for p in range(0, len(data), block):
        chunk = data[p:p + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        body = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(body) + 8 - 1
        assert bsize < 65536, "incompressible chunk: use a smaller block"
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + body +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    if eof:
        out.append(EOF_MARKER)
    return b"".join(out)
