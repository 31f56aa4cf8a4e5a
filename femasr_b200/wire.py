"""Compact wire format for codebook indices: a [B,1,h,w] index map against an n_e-entry codebook
is ceil(log2(n_e)) bits per code (10 bits for the shipped 1024-entry codebook) - 1.25 bytes stand for an 8x8x3 fp32
output patch of `FeMaSRNet.decode_indices` (768 bytes), against 8 bytes for the int64 tensor the reference passes
around (`femasr_arch.py:100,376-385`; `vis_codebook.py:56-83`).  Pure tensor arithmetic: works on any device, so the
codes can be packed on the GPU before the D2H copy and unpacked on the GPU right before `decode_indices`."""
from __future__ import annotations

from typing import Tuple

import torch


def code_bits(n_e: int) -> int:
    if n_e < 2:
        raise ValueError("a codebook needs at least two entries")
    return (int(n_e) - 1).bit_length()


def packed_nbytes(numel: int, n_e: int) -> int:
    return (numel * code_bits(n_e) + 7) // 8


def _lib():
    from . import lib as L
    return L


def pack_codes(indices: torch.Tensor, n_e: int) -> torch.Tensor:
    """[...] integer codes in [0, n_e) -> uint8[packed_nbytes]; little-endian bit order (code i occupies bits
    [i*w, (i+1)*w) of the stream, least significant bit first).  CUDA tensors are packed by a device kernel
    (`femasr_pack_codes`: one thread per eight codes), CPU tensors by the tensor arithmetic below (same stream)."""
    w = code_bits(n_e)
    flat = indices.reshape(-1).to(torch.int64)
    if flat.is_cuda and flat.numel() and n_e <= 65536:
        L = _lib()
        flat = flat.contiguous()
        with torch.cuda.device(flat.device):
            out = torch.empty(packed_nbytes(flat.numel(), n_e), dtype=torch.uint8, device=flat.device)
            status = torch.zeros(1, dtype=torch.int32, device=flat.device)
            L.check(L.load().femasr_pack_codes(flat.data_ptr(), out.data_ptr(), flat.numel(), int(n_e), status.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
            if int(status.item()):
                raise ValueError(f"codes must lie in [0, {n_e})")
        return out
    if flat.numel() and (int(flat.min()) < 0 or int(flat.max()) >= n_e):
        raise ValueError(f"codes must lie in [0, {n_e})")
    shifts = torch.arange(w, device=flat.device, dtype=torch.int64)
    bits = ((flat[:, None] >> shifts) & 1).to(torch.uint8).reshape(-1)          # [N*w] stream, LSB first
    pad = (-bits.numel()) % 8
    if pad:
        bits = torch.cat([bits, bits.new_zeros(pad)])
    weights = (1 << torch.arange(8, device=flat.device, dtype=torch.int64)).to(torch.uint8)
    return (bits.view(-1, 8) * weights).sum(1, dtype=torch.int64).to(torch.uint8)


def unpack_codes(packed: torch.Tensor, shape: Tuple[int, ...], n_e: int) -> torch.Tensor:
    """Inverse of pack_codes: uint8 stream -> int64 tensor of `shape` (e.g. [B,1,h,w] for decode_indices)."""
    w = code_bits(n_e)
    numel = 1
    for d in shape:
        numel *= int(d)
    if packed.dtype != torch.uint8 or packed.numel() != packed_nbytes(numel, n_e):
        raise ValueError(f"expected {packed_nbytes(numel, n_e)} uint8 values for {numel} codes of {w} bits")
    if packed.is_cuda and numel and n_e <= 65536:
        L = _lib()
        packed = packed.contiguous()
        with torch.cuda.device(packed.device):
            out = torch.empty(numel, dtype=torch.int64, device=packed.device)
            L.check(L.load().femasr_unpack_codes(packed.data_ptr(), out.data_ptr(), numel, int(n_e),
                                                 torch.cuda.current_stream().cuda_stream))
        return out.view(*shape)
    shifts = torch.arange(8, device=packed.device, dtype=torch.int64)
    bits = ((packed.to(torch.int64)[:, None] >> shifts) & 1).reshape(-1)[: numel * w].view(numel, w)
    vals = (bits << torch.arange(w, device=packed.device, dtype=torch.int64)).sum(1)
    return vals.view(*shape)
