#!/usr/bin/env python
"""Per-layer timing of the C1 SoundStream encode (32 clips x 2 s @ 24 kHz): CUDA events around every launch."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200 import ops  # noqa: E402
from audiolm_pytorch_b200.soundstream import SoundStream  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(7)
ss = SoundStream(codebook_size=1024, rq_num_quantizers=8, target_sample_hz=24000, use_local_attn=False)
for layer in ss.rq.rvqs[0].layers:
    layer._codebook.embed.normal_()
    layer._codebook.initted.fill_(True)
ss = ss.cuda().eval()
wave = torch.randn(clips, 48000, device="cuda")
with torch.no_grad():
    for _ in range(3):
        ss(wave, return_encoded=True)
    ops.PROFILE_SHAPES = True
    ops.profile_start()
    for _ in range(3):
        ss(wave, return_encoded=True)
    prof = ops.profile_stop()
rows = []
for cls, (ms, work, n) in prof.items():
    unit = ops.CLASS_UNIT.get(cls)
    rate = work / (ms * 1e-3) / (1e9 if unit == "byte" else 1e12)
    rows.append((cls, ms / 3, n // 3, rate, "GB/s" if unit == "byte" else "TFLOP/s"))
for r in rows:
    print(f"{r[0]:55s} {r[1]*1e3:9.1f} us  x{r[2]}  {r[3]:9.1f} {r[4]}")
print(json.dumps({"total_us": sum(r[1] for r in rows) * 1e3}))
