timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_properties_gpu.py::test_full_size_logits_vs_cpu_oracle 2>&1 | tail -3
timeout 600 python tools/bench_decode.py 512 > gpurun_out/decode_latency_v3.json 2>gpurun_out/decode_latency_v3.err; cat gpurun_out/decode_latency_v3.json
