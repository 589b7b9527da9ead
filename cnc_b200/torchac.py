"""torchac-shaped entry points backed by the GPU range coder (cnc_b200/csrc/coder.cu).

Mirrors the two functions the reference calls (examples/utils_bpp_acc.py:87,108):
    encode_float_cdf(cdf_float [N,3], sym int16 [N], check_input_bounds=...) -> bytes
    decode_float_cdf(cdf_float [N,3], byte_stream) -> int16 [N]
for the binary alphabet CNC uses (Lp == 3, cdf = [0, 1-p, 1]).  The batched
`encode_streams / decode_streams` are what the codec driver uses: all streams of a
bitstream set are coded by one kernel launch, one warp per stream.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from ._lib import check, lib, ptr, stream


def cdf_from_p(p: torch.Tensor) -> torch.Tensor:
    """uint16-valued (stored as int16 bit pattern in a torch.int16 tensor) c1 = round((1-p)*65534)+1."""
    p = p.contiguous().float().view(-1)
    c1 = torch.empty(p.numel(), dtype=torch.int16, device=p.device)
    check(lib().cnc_cdf_from_p(ptr(p), ptr(c1), p.numel(), stream()))
    return c1


def _offsets(lengths: Sequence[int], align: int = 1):
    off = [0]
    for n in lengths:
        off.append(off[-1] + (int(n) + align - 1) // align * align)
    return off


class _EncodeJob:
    """Streams handed to the GPU coder whose bytes have not been fetched yet (`encode_streams_async`)."""

    def __init__(self, c1, sym, lens, caps, out, out_off_h, out_len, cuda_stream):
        self.c1, self.sym, self.lens, self.caps = c1, sym, lens, caps
        self.out, self.out_off_h, self.out_len, self.cuda_stream = out, out_off_h, out_len, cuda_stream

    def result(self) -> List[bytes]:
        """Wait for the coder and return the K byte strings."""
        self.cuda_stream.synchronize()
        lens_h = self.out_len.cpu().tolist()
        if any(l > c for l, c in zip(lens_h, self.caps)):   # pathological probabilities: redo with larger buffers
            caps = [max(c, l + 64) for l, c in zip(lens_h, self.caps)]
            with torch.cuda.stream(self.cuda_stream):
                job = _launch_encode(self.c1, self.sym, self.lens, caps)
            return job.result()
        host = self.out.cpu().numpy()
        return [host[self.out_off_h[k]: self.out_off_h[k] + lens_h[k]].tobytes() for k in range(len(lens_h))]


def _launch_encode(c1, sym, lens, caps) -> _EncodeJob:
    dev = c1.device
    K = len(lens)
    sym_off = torch.tensor(_offsets(lens), dtype=torch.int64, device=dev)
    out_off_h = _offsets(caps, 4)
    out = torch.empty(out_off_h[-1], dtype=torch.uint8, device=dev)
    out_off = torch.tensor(out_off_h, dtype=torch.int64, device=dev)
    out_len = torch.zeros(K, dtype=torch.int64, device=dev)
    check(lib().cnc_ac_encode(ptr(c1), ptr(sym), ptr(sym_off), ptr(out), ptr(out_off), ptr(out_len), K, stream()))
    job = _EncodeJob(c1, sym, lens, caps, out, out_off_h, out_len, torch.cuda.current_stream(dev))
    job._keep = (sym_off, out_off)
    return job


def encode_streams_async(c1_list: Sequence[torch.Tensor], sym_list: Sequence[torch.Tensor]) -> _EncodeJob:
    """Launch the coder for K independent streams on the current CUDA stream without waiting for it;
    `.result()` returns the byte strings.  Lets a caller code early streams while it still computes the
    probabilities of later ones (the coder occupies one SM per stream)."""
    K = len(c1_list)
    if K == 0:
        raise ValueError("no streams")
    lens = [int(c.numel()) for c in c1_list]
    c1 = (torch.cat([c.view(-1) for c in c1_list]) if K > 1 else c1_list[0].view(-1)).contiguous()
    sym = (torch.cat([x.view(-1) for x in sym_list]) if K > 1 else sym_list[0].view(-1)).to(torch.uint8).contiguous()
    return _launch_encode(c1, sym, lens, [n // 8 * 2 + 64 for n in lens])


def encode_streams(c1_list: Sequence[torch.Tensor], sym_list: Sequence[torch.Tensor]) -> List[bytes]:
    """Encode K independent streams (c1 int16-bit-pattern, sym uint8 in {0,1}) -> K byte strings."""
    if len(c1_list) == 0:
        return []
    return encode_streams_async(c1_list, sym_list).result()


def decode_streams(c1_list: Sequence[torch.Tensor], streams: Sequence[bytes]) -> List[torch.Tensor]:
    """Decode K independent streams -> K uint8 symbol tensors (on the device of c1)."""
    K = len(c1_list)
    if K == 0:
        return []
    dev = c1_list[0].device
    lens = [int(c.numel()) for c in c1_list]
    c1 = (torch.cat([c.view(-1) for c in c1_list]) if K > 1 else c1_list[0].view(-1)).contiguous()
    sym_off_h = _offsets(lens)
    sym_off = torch.tensor(sym_off_h, dtype=torch.int64, device=dev)
    in_off_h = _offsets([len(b) for b in streams], 4)
    import numpy as np

    buf = np.zeros(max(in_off_h[-1], 4), np.uint8)
    for k, b in enumerate(streams):
        buf[in_off_h[k]: in_off_h[k] + len(b)] = np.frombuffer(b, np.uint8)
    inp = torch.from_numpy(buf).to(dev)
    in_off = torch.tensor(in_off_h[:-1], dtype=torch.int64, device=dev)
    in_len = torch.tensor([len(b) for b in streams], dtype=torch.int64, device=dev)
    sym = torch.empty(max(sym_off_h[-1], 1), dtype=torch.uint8, device=dev)
    check(lib().cnc_ac_decode(ptr(c1), ptr(sym_off), ptr(inp), ptr(in_off), ptr(in_len), ptr(sym), K, stream()))
    return [sym[sym_off_h[k]: sym_off_h[k + 1]] for k in range(K)]


def _binary_c1(cdf_float: torch.Tensor) -> torch.Tensor:
    if cdf_float.shape[-1] != 3:
        raise ValueError("cnc_b200.torchac only implements the binary alphabet used by CNC (Lp == 3)")
    if not cdf_float.is_cuda:
        raise RuntimeError("cdf_float must be a CUDA tensor (the coder runs on the GPU; there is no CPU path)")
    # torchac: cdf_int = round(cdf_float * (2^16 - (Lp-1))) + arange(Lp); column 1 is the only free entry
    col = cdf_float[..., 1].contiguous().float().view(-1)
    # cdf_float[...,1] == 1 - p was formed by the caller; feed p' = 1 - col would re-round, so
    # quantise the column directly with the same formula (1 - (1 - col) is exact only sometimes).
    v = torch.round(col * 65534.0).to(torch.int32) + 1
    return v.to(torch.int16)


def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False) -> bytes:
    if check_input_bounds:
        if cdf_float.min() < 0 or cdf_float.max() > 1:
            raise ValueError("cdf_float out of [0, 1]")
        if sym.max() >= cdf_float.shape[-1] - 1:
            raise ValueError("sym out of range")
    c1 = _binary_c1(cdf_float)
    return encode_streams([c1], [sym.to(c1.device).view(-1).to(torch.uint8)])[0]


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
    c1 = _binary_c1(cdf_float)
    return decode_streams([c1], [byte_stream])[0].to(torch.int16).view(cdf_float.shape[:-1])
