"""CPU: the compact index wire format (femasr_b200/wire.py) round-trips and has the advertised size."""
import pytest
import torch

from femasr_b200.wire import code_bits, pack_codes, packed_nbytes, unpack_codes


@pytest.mark.parametrize("n_e,shape", [(1024, (2, 1, 16, 24)), (1024, (1, 1, 3, 5)), (512, (3, 1, 7, 9)), (256, (1, 1, 64, 64)),
                                       (1000, (1, 1, 4, 4)), (2, (1, 1, 5, 3))])
def test_roundtrip_and_size(n_e, shape):
    g = torch.Generator().manual_seed(n_e + shape[2])
    idx = torch.randint(0, n_e, shape, generator=g)
    idx.view(-1)[0] = n_e - 1                       # the extreme codes survive
    idx.view(-1)[-1] = 0
    p = pack_codes(idx, n_e)
    assert p.dtype == torch.uint8 and p.numel() == packed_nbytes(idx.numel(), n_e) == (idx.numel() * code_bits(n_e) + 7) // 8
    back = unpack_codes(p, shape, n_e)
    assert back.dtype == torch.int64 and torch.equal(back, idx)


def test_ten_bits_for_the_shipped_codebook():
    assert code_bits(1024) == 10 and code_bits(512) == 9 and code_bits(1025) == 11
    # 64x64 codes of a 512x512 output: 5120 bytes against 32768 for int64 and 3 MiB for the fp32 image they decode to
    assert packed_nbytes(64 * 64, 1024) == 5120


def test_bad_input_is_rejected():
    with pytest.raises(ValueError):
        pack_codes(torch.tensor([0, 1024]), 1024)
    with pytest.raises(ValueError):
        unpack_codes(torch.zeros(3, dtype=torch.uint8), (1, 1, 2, 2), 1024)
