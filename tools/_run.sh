timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_models_gpu.py tests/test_properties_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/profile_step.py 2>/dev/null | grep "mnmn\|step breakdown" | head -12
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r27.json 2> gpurun_out/bench_r27.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r27.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['kernels']['gemm_bf16_tcgen05'], d['kernels']['gemm_skinny_hc_param_grad'])"
