timeout 600 python -m pytest tests/test_me_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/ab_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_def.json 2> gpurun_out/ab_def.err
for v in t128b8 t128b7 t256b5; do
B200RDO_LIB=$PWD/rav1e_b200/variants/libb200rdo_$v.so timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
B200RDO_LIB=$PWD/rav1e_b200/variants/libb200rdo_t128b8.so timeout 600 python -m pytest tests/test_me_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/ab_tests_t128.log
