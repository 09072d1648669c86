"""Deterministic, device-independent test tensors.

Restates test::seeded_tensor (tests/core/layers/mlu/tests_utils.cpp:159-274):
FNV-1a-64(key) seeds a SplitMix64 stream; floats are (u >> 11) * 2^-53 in
[0, 1) computed in double and cast to the target dtype.  SplitMix64 is
counter-based (state_i = seed + (i+1)*GAMMA), so the stream is vectorised.
"""
import numpy as np
import torch

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def fnv1a64(key: str) -> int:
    h = 0xCBF29CE484222325
    for c in key.encode("utf-8"):
        h ^= c
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def splitmix64_stream(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * _GAMMA
        z ^= z >> np.uint64(30)
        z *= _C1
        z ^= z >> np.uint64(27)
        z *= _C2
        z ^= z >> np.uint64(31)
    return z


def seeded_tensor(key: str, shape, dtype=torch.bfloat16) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = splitmix64_stream(fnv1a64(key), n)
    if dtype.is_floating_point:
        vals = (u >> np.uint64(11)).astype(np.float64) * (1.0 / float(1 << 53))
        return torch.from_numpy(vals).to(dtype).view(*shape).contiguous()
    if dtype == torch.bool:
        return torch.from_numpy((u & np.uint64(1)).astype(np.bool_)).view(*shape)
    info = torch.iinfo(dtype)
    span = info.max - info.min + 1
    vals = [info.min + (int(x) % span) for x in u.tolist()]
    return torch.tensor(vals, dtype=dtype).view(*shape)
