"""TEST INFRASTRUCTURE -- POSIX `cksum` (CRC-32/CKSUM, polynomial 0x04C11DB7, length appended).

The reference's only tests compare `cksum` output of each tool's stdout against
hard-coded values (test_data/test_suite.py:7-27); this reproduces that checksum
in-process so the goldens can be checked without a shell.
"""

_TABLE = []
for _i in range(256):
    _c = _i << 24
    for _ in range(8):
        _c = ((_c << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if _c & 0x80000000 else (_c << 1) & 0xFFFFFFFF
    _TABLE.append(_c)


def cksum(data):
    """Return (crc, nbytes) exactly as `cksum file` prints them."""
    if isinstance(data, str):
        data = data.encode("latin-1")
    crc = 0
    tbl = _TABLE
    for b in data:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ tbl[(crc >> 24) ^ b]
    n = len(data)
    while n:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ tbl[(crc >> 24) ^ (n & 0xFF)]
        n >>= 8
    return (~crc) & 0xFFFFFFFF, len(data)
