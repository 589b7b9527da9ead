"""Single-file container for a compressed CNC scene.

The reference writes the 33 entropy-coded streams as loose `<prefix>_*.b` files and keeps everything else the decoder
needs in Python variables of the training process (`Pgs_dict`, the occupancy grid `estimator.binaries`, the MLP
weights), accounting for their size with two estimates:

    quantize_params(dict, digits)      examples/train_CNC_nerf_synthetic.py:30-50   (MLP weights, uniform `digits`-bit grid)
    get_binary_vxl_size(binary_vxl)    examples/train_CNC_nerf_synthetic.py:53-68   (occupancy, zeroth-order entropy)

Both are mirrored here with the same arithmetic and return values, and `pack` / `unpack` turn the estimate into bytes:
one blob = header (json: layout, stream index, tensor index) + per-level frequencies `Pgs_dict` as raw fp32 (they feed
the CDFs, so they must survive bit for bit) + occupancy grid at 1 bit per cell + MLP weights as packed `digits`-bit
integers with their (min, interval) pair + the streams back to back.  `decode_binary_vxl_mixPg_3D2D(..., streams=...)`
consumes `unpack(...)["streams"]` directly.  Pure host-side byte shuffling: numpy + torch CPU ops, no kernel.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Optional

import numpy as np
import torch

MAGIC = b"CNCB200\x01"


def quantize_params(dict_input: Dict[str, torch.Tensor], ss=None, digits: int = 10):
    """train_CNC_nerf_synthetic.py:30-50: (MB quantised, MB fp32, {name: dequantised tensor}, [quantised values])"""
    bits, bits_orig = 0, 0
    dict_quantized, quantized_v_list = {}, []
    for n, p_input in dict_input.items():
        min_v, max_v = torch.min(p_input), torch.max(p_input)
        scales = 2 ** digits - 1
        interval = (max_v - min_v) / scales + 1e-6   # avoid 0 if max_v == min_v
        quantized_v = (p_input - min_v) // interval
        dict_quantized[n] = quantized_v * interval + min_v
        quantized_v_list.append(quantized_v)
        bits += digits * p_input.numel() + 32 + 32   # + min_v and scale
        bits_orig += 32 * p_input.numel()
    return bits / 8.0 / 1024 / 1024, bits_orig / 8.0 / 1024 / 1024, dict_quantized, quantized_v_list


def get_binary_vxl_size(binary_vxl: torch.Tensor):
    """train_CNC_nerf_synthetic.py:53-68: (P(occupied), zeroth-order entropy in MB incl. 32 bits for Pg, cell count)"""
    with torch.no_grad():
        ttl_num = binary_vxl.numel()
        pos_num = torch.sum(binary_vxl)
        neg_num = ttl_num - pos_num
        Pg = pos_num / ttl_num
        ttl_bit = pos_num * (-torch.log2(Pg)) + neg_num * (-torch.log2(1 - Pg)) + 32
    return Pg, ttl_bit.item() / 8.0 / 1024 / 1024, ttl_num


def _pack_uint(values: np.ndarray, digits: int) -> bytes:
    """little-endian bit packing of non-negative integers < 2^digits"""
    v = values.astype(np.uint64).reshape(-1)
    bits = ((v[:, None] >> np.arange(digits, dtype=np.uint64)[None, :]) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(bits, bitorder="little").tobytes()


def _unpack_uint(data: bytes, n: int, digits: int) -> np.ndarray:
    bits = np.unpackbits(np.frombuffer(data, np.uint8), bitorder="little")[: n * digits].reshape(n, digits).astype(np.uint64)
    return (bits << np.arange(digits, dtype=np.uint64)[None, :]).sum(1)


def quantize_state(state: Dict[str, torch.Tensor], digits: int = 13) -> Dict[str, tuple]:
    """{name: (integer levels [shape] int64, min fp32, interval fp32)}: the `digits`-bit grid of quantize_params, kept as
    integers so that what `pack` stores is exactly what the encoder side used (re-quantising a dequantised tensor does not
    reproduce it: the interval is derived from max - min, which shrinks)"""
    out = {}
    for name, t in state.items():
        t32 = t.detach().float().cpu()
        min_v, max_v = torch.min(t32), torch.max(t32)
        interval = (max_v - min_v) / (2 ** digits - 1) + 1e-6
        out[name] = (((t32 - min_v) // interval).to(torch.int64), min_v.clone(), interval.clone())
    return out


def dequantize_state(qstate: Dict[str, tuple], device="cpu") -> Dict[str, torch.Tensor]:
    """the tensors a decoder gets back from `unpack` for this quantised state, bit for bit (same fp32 arithmetic on the CPU).
    An encoder must run the context models with THESE weights (train_CNC_nerf_synthetic.py:513-520 only accounts for the
    13 bits; a stand-alone decoder actually has nothing else)."""
    return {k: (q.to(torch.float32) * i.to(torch.float32) + m.to(torch.float32)).to(device) for k, (q, m, i) in qstate.items()}


def pack(streams: Dict[str, bytes], Pgs_dict: Dict[str, torch.Tensor], binary_vxl: torch.Tensor,
         mlp_state: Optional[Dict[str, torch.Tensor]] = None, layout: Optional[dict] = None, digits: int = 13) -> bytes:
    """-> one self-contained blob.  `streams` as returned by encode_binary_vxl_mixPg_3D2D(..., return_streams=True);
    `mlp_state`: every non-table tensor the decoder side needs (field MLPs, context models), as tensors (quantised here)
    or as the triples of `quantize_state` (stored as they are); `layout`: free-form json (resolutions, hash sizes, the
    symbol-order seed: `CNC_context_models.layout()`) echoed back by `unpack`."""
    sections, index = [], {"layout": layout or {}, "digits": digits, "streams": [], "pgs": [], "tensors": []}

    def add(b: bytes) -> int:
        sections.append(b)
        return len(b)

    names = list(Pgs_dict.keys())
    pg = np.array([float(Pgs_dict[k].detach().float().cpu()) for k in names], np.float32)
    index["pgs"] = names
    add(pg.tobytes())
    vx = binary_vxl.detach().cpu().numpy().astype(bool)
    index["occupancy_shape"] = list(vx.shape)
    add(np.packbits(vx.reshape(-1), bitorder="little").tobytes())
    mlp_state = mlp_state or {}
    raw = {k: v for k, v in mlp_state.items() if isinstance(v, torch.Tensor)}
    qstate = {**quantize_state(raw, digits), **{k: v for k, v in mlp_state.items() if not isinstance(v, torch.Tensor)}}
    for name in mlp_state:
        q, min_v, interval = qstate[name]
        index["tensors"].append({"name": name, "shape": list(q.shape), "bytes": add(
            struct.pack("<ff", float(min_v), float(interval)) + _pack_uint(q.numpy(), digits))})
    for name, data in streams.items():
        index["streams"].append({"name": name, "bytes": add(bytes(data))})
    head = json.dumps(index, separators=(",", ":")).encode()
    return MAGIC + struct.pack("<I", len(head)) + head + b"".join(sections)


def unpack(blob: bytes, device="cpu") -> dict:
    """inverse of `pack`: {"layout", "Pgs_dict", "binary_vxl", "mlp_state" (dequantised like quantize_params), "streams"}"""
    if blob[:8] != MAGIC:
        raise ValueError("not a cnc-b200 container")
    (hl,) = struct.unpack("<I", blob[8:12])
    index = json.loads(blob[12:12 + hl].decode())
    pos = 12 + hl

    def take(n: int) -> bytes:
        nonlocal pos
        b = blob[pos:pos + n]
        if len(b) != n:
            raise ValueError("truncated container")
        pos += n
        return b

    digits = index["digits"]
    pg = np.frombuffer(take(4 * len(index["pgs"])), np.float32)
    Pgs = {k: torch.tensor(pg[i], dtype=torch.float32, device=device) for i, k in enumerate(index["pgs"])}
    shape = index["occupancy_shape"]
    ncell = int(np.prod(shape))
    vx = np.unpackbits(np.frombuffer(take((ncell + 7) // 8), np.uint8), bitorder="little")[:ncell].astype(bool).reshape(shape)
    mlp = {}
    for t in index["tensors"]:
        raw = take(t["bytes"])
        min_v, interval = struct.unpack("<ff", raw[:8])
        n = int(np.prod(t["shape"])) if t["shape"] else 1
        q = torch.from_numpy(_unpack_uint(raw[8:], n, digits).astype(np.float32)).reshape(t["shape"])
        mlp[t["name"]] = (q * torch.tensor(interval, dtype=torch.float32) + torch.tensor(min_v, dtype=torch.float32)).to(device)
    streams = {s["name"]: take(s["bytes"]) for s in index["streams"]}
    return {"layout": index["layout"], "Pgs_dict": Pgs, "binary_vxl": torch.from_numpy(vx).to(device), "mlp_state": mlp,
            "streams": streams}
