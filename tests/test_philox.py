"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the oracle's and the product's generator."""
import gameai_b200 as g

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_oracle_philox_kat(oracle):
    for ctr, key, out in KAT:
        assert oracle.philox(ctr, key) == out


def test_product_philox_kat():
    for ctr, key, out in KAT:
        assert g.philox4x32_10(ctr, key) == out


def test_product_matches_oracle_stream(oracle):
    import random
    r = random.Random(3)
    for _ in range(200):
        ctr = tuple(r.getrandbits(32) for _ in range(4)); key = tuple(r.getrandbits(32) for _ in range(2))
        assert g.philox4x32_10(ctr, key) == oracle.philox(ctr, key)
