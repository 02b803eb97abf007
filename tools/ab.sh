python -m pytest tests/test_me_gpu.py tests/test_txfm_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/ab_p32.json 2> gpurun_out/ab_p32.err
python bench.py --steps 10 --warmup 3 --pairs-per-launch 1 > gpurun_out/ab_p1.json 2> gpurun_out/ab_p1.err
python bench.py --steps 10 --warmup 3 --pairs-per-launch 8 > gpurun_out/ab_p8.json 2> gpurun_out/ab_p8.err
B200RDO_LIB=$PWD/rav1e_b200/variants/libb200rdo_mb4.so python bench.py --steps 10 --warmup 3 > gpurun_out/ab_mb4_p32.json 2> gpurun_out/ab_mb4_p32.err
B200RDO_LIB=$PWD/rav1e_b200/variants/libb200rdo_mb4.so python bench.py --steps 10 --warmup 3 --pairs-per-launch 1 > gpurun_out/ab_mb4_p1.json 2> gpurun_out/ab_mb4_p1.err
