timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_properties_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
r = bench.codec_encode_bench(torch.device('cuda:0'))
print(r['value'], r['ms_per_call'], r['kernels'])
PY
