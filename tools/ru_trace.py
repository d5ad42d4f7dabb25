#!/usr/bin/env python
"""Wait-time breakdown of the ResidualUnit kernel (needs libalm_b200.so built with ALM_EXTRA_NVCC_FLAGS=-DALM_RU_TRACE)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
buf = (C.c_ulonglong * 16)()
names = ["mma:a_full", "mma:a2_full", "mma:d2_empty", "mma:w_full", "epi0:d1_full", "epi0:d2_full"]
for Cc, T in ((32, 48000), (64, 24000), (128, 6000), (256, 1200)):
    for d in (1, 9):
        x = ops.c8s_pack(torch.randn(32, Cc, T, device="cuda"))
        wu = ops.pack_ru_weights(torch.randn(Cc, Cc, 7, device="cuda") * 0.05, torch.randn(Cc, Cc, 1, device="cuda") * 0.1)
        b = torch.zeros(Cc, device="cuda")
        for _ in range(2):
            ops.codec_ru_tc(x, wu, b, b, dilation=d)
        lib.alm_debug_ru_trace(buf, 1)
        ops.codec_ru_tc(x, wu, b, b, dilation=d)
        lib.alm_debug_ru_trace(buf, 1)
        v = list(buf)
        ctas = max(v[9], 1)
        tot = v[8] / ctas
        tiles = 32 * -(-T // 128) / ctas
        print(f"C{Cc} d{d}: MMA-thread lifetime {tot:9.0f} clk/CTA ({tot / tiles:7.0f} per tile, {tiles:.1f} tiles/CTA, {ctas} CTAs) | " +
              " ".join(f"{n} {v[i] / ctas / tot * 100:4.1f}%" for i, n in enumerate(names)))
