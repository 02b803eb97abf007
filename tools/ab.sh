timeout 600 python -m pytest tests/test_quantize_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab_tests.log
