timeout 600 python -m pytest tests/test_me_search_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/ab_tests.log
