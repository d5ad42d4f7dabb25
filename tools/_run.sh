timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_properties_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
from audiolm_pytorch_b200 import ops, soundstream
from audiolm_pytorch_b200.soundstream import SoundStream
r = bench.codec_encode_bench(torch.device('cuda:0'))
print(r['value'], r['ms_per_call'], r['kernels'])
# bit-identity fused vs unfused
torch.manual_seed(7)
dev = torch.device('cuda:0')
ss = SoundStream(codebook_size=1024, rq_num_quantizers=8, target_sample_hz=24000, use_local_attn=False)
for rvq in ss.rq.rvqs:
    for layer in rvq.layers:
        layer._codebook.embed.normal_(); layer._codebook.initted.fill_(True)
ss = ss.to(dev).eval()
wave = torch.randn(4, 48000, device=dev)
with torch.no_grad():
    a = ss(wave, return_encoded=True)
    soundstream.FUSE_RESIDUAL_UNITS = False
    b = ss(wave, return_encoded=True)
    soundstream.FUSE_RESIDUAL_UNITS = True
    ops.PROFILE_SHAPES = True
    ops.profile_start(); ss(torch.randn(32, 48000, device=dev), return_encoded=True); prof = ops.profile_stop()
print('fused == unfused: quantized', torch.equal(a[0], b[0]), 'indices', torch.equal(a[1], b[1]))
for cls, (ms, work, n) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"{cls:60s} n={n} {ms:8.3f} ms {work/ms/1e9:7.2f} TFLOP/s")
PY
