"""The operand-image format of the tensor-core training path (include/inrf.h, csrc/common.cuh IMG_*), pinned on the CPU:
element (row, col) of a 128 x 64 fp16 chunk lives at byte (row>>3)*1024 + (row&7)*128 + (((col>>3) ^ (row&7)) << 4) +
(col&7)*2 - the UMMA SWIZZLE_128B atom, which the kernels read K-major (rows = M/N) and MN-major (rows = K)."""
import torch

from tests.util import decode_images, encode_images


def _offset(row, col):
    return (row >> 3) * 1024 + (row & 7) * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2


def test_round_trip_and_byte_offsets():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 128, 64, generator=g).to(torch.float16).float()
    buf = encode_images(x)
    assert buf.numel() == 3 * 5 * 16384
    assert torch.equal(decode_images(buf, 3, 5), x)
    halves = buf.view(torch.float16)
    for t, s, row, col in ((0, 0, 0, 0), (0, 0, 1, 0), (0, 0, 7, 63), (1, 3, 8, 8), (2, 4, 127, 63), (2, 1, 77, 29)):
        off = (t * 5 + s) * 16384 + _offset(row, col)
        assert off % 2 == 0 and float(halves[off // 2]) == float(x[t, s, row, col]), (t, s, row, col)


def test_offsets_are_a_bijection_with_conflict_free_columns():
    offs = {_offset(r, c) for r in range(128) for c in range(64)}
    assert len(offs) == 128 * 64 and min(offs) == 0 and max(offs) == 16384 - 2
    # the 8 rows of an atom put the same logical 16-byte unit into 8 different bank groups (the point of the swizzle)
    for u in range(8):
        assert len({(_offset(r, 8 * u) >> 4) & 7 for r in range(8)}) == 8
