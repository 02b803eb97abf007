timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_up.json 2> gpurun_out/ab_up.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_up2.json 2> gpurun_out/ab_up2.err
