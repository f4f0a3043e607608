"""CPU emulation of rows_write_kernel's ordinal arithmetic (csrc/select_kernels.cuh): lane l holds rows
s*128 + l*4 + j of a 512-row chunk as mask bit s*4+j; per-sub-chunk counts are packed 8 bits each and scanned
across lanes; the ordinal of a passing row must equal its rank among the chunk's passing rows in ROW order."""
import numpy as np


def emulate(pass_rows):
    masks = np.zeros(32, dtype=np.uint32)
    for r in pass_rows:
        s, rem = divmod(r, 128)
        l, j = divmod(rem, 4)
        masks[l] |= np.uint32(1 << (s * 4 + j))
    popc = lambda x: bin(int(x)).count("1")
    packed = np.array([popc(m & 0xF) | (popc(m & 0xF0) << 8) | (popc(m & 0xF00) << 16) | (popc(m & 0xF000) << 24)
                       for m in masks], dtype=np.uint64)
    incl = np.cumsum(packed)
    assert int(incl[31]) < 2**32 and all(((int(incl[31]) >> (8 * s)) & 0xFF) <= 128 for s in range(4))
    tot = int(incl[31])
    out = {}
    for l in range(32):
        excl = int(incl[l] - packed[l])
        base = 0
        for s in range(4):
            o = base + ((excl >> (8 * s)) & 0xFF)
            for j in range(4):
                if int(masks[l]) & (1 << (s * 4 + j)):
                    out[s * 128 + l * 4 + j] = o
                    o += 1
            base += (tot >> (8 * s)) & 0xFF
    return out


def test_ordinals_are_row_order_ranks():
    rng = np.random.default_rng(3)
    for density in (0.0, 0.01, 0.2, 0.5, 1.0):
        for _ in range(20):
            rows = np.nonzero(rng.random(512) < density)[0].tolist()
            got = emulate(rows)
            assert got == {r: i for i, r in enumerate(sorted(rows))}
