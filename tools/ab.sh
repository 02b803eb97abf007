timeout 600 python -m pytest tests/test_me_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_satd.json 2> gpurun_out/ab_satd.err
