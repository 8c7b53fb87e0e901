"""Host restatements of the GROUP4 and SAMPLE16 panel layouts (simu_b200.plink.interleave_*): the contract of
include/simu_b200.h, checked here without a GPU; tests/test_gpu_parity.py compares the device image with it."""
import numpy as np

from oracle import oracle
from simu_b200 import plink


def test_group4_image_is_the_four_snp_table_index():
    n, m = 203, 11
    rows = oracle.synth_rows(7, 0, m, n)
    img = plink.interleave_group4(rows, n)
    assert img.shape == (3, 256)
    codes = np.stack([oracle.unpack(rows[j], n) for j in range(m)])          # libplinkio codes 0,1,2,3
    bits = np.array([0b00, 0b10, 0b11, 0b01], np.uint8)[codes]                # back to the packed bit pairs (bed.c:75-85)
    for g in range(3):
        want = np.zeros(n, np.uint8)
        for r in range(4):
            if 4 * g + r < m:
                want |= bits[4 * g + r] << (2 * r)
        assert np.array_equal(img[g, :n], want)
        assert not img[g, n:].any()                                            # pitch padding stays 00


def test_group4_round_trip_and_known_answer_bytes():
    n, m = 1001, 19
    rows = oracle.synth_rows(3, 5, m, n)
    B = plink.row_bytes(n)
    back = plink.deinterleave_group4(plink.interleave_group4(rows, n), m, n)
    masked = rows[:, :B].copy()
    if n % 4:
        masked[:, B - 1] &= (1 << (2 * (n % 4))) - 1                           # padding pairs are not carried over
    assert np.array_equal(back, masked)
    # bed_test.c:237-250: byte 0x78 = fields 00,10,11,01 for samples 0..3
    img = plink.interleave_group4(np.array([[0x78], [0x2D]], np.uint8), 4)
    assert img[0, :4].tolist() == [0b0100, 0b1110, 0b1011, 0b0001]


def test_sample16_image_is_the_dosage_coded_sample_word():
    n, m = 203, 37
    rows = oracle.synth_rows(7, 0, m, n)
    img = plink.interleave_sample16(rows, n)
    assert img.shape == (3, 256) and img.dtype == np.uint32
    codes = np.stack([oracle.unpack(rows[j], n) for j in range(m)]).astype(np.uint32)   # libplinkio codes: 0, 1, 2 = A2 dosage, 3 = missing
    for g in range(3):
        want = np.zeros(n, np.uint32)
        for r in range(16):
            if 16 * g + r < m:
                want |= codes[16 * g + r] << np.uint32(2 * r)
        assert np.array_equal(img[g, :n], want)
        assert not img[g, n:].any()                                            # pitch padding stays 0


def test_sample16_round_trip_and_known_answer_bytes():
    n, m = 1001, 19
    rows = oracle.synth_rows(3, 5, m, n)
    B = plink.row_bytes(n)
    back = plink.deinterleave_sample16(plink.interleave_sample16(rows, n), m, n)
    masked = rows[:, :B].copy()
    if n % 4:
        masked[:, B - 1] &= (1 << (2 * (n % 4))) - 1                           # padding pairs are not carried over
    assert np.array_equal(back, masked)
    # bed_test.c:237-250: byte 0x78 = codes 0,1,2,3 for samples 0..3; 0x2d = 3,2,1,0: row 0 in bits 0-1, row 1 in bits 2-3
    img = plink.interleave_sample16(np.array([[0x78], [0x2D]], np.uint8), 4)
    assert img[0, :4].tolist() == [0 | 3 << 2, 1 | 2 << 2, 2 | 1 << 2, 3 | 0 << 2]
