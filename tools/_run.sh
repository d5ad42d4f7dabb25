timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_properties_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py hc
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r28.json 2> gpurun_out/bench_r28.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r28.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac']);
[print(k, v) for k, v in d['kernels'].items() if 'hc' in k]; print(d['soundstream_encode']['value'])"
