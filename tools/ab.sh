timeout 600 python -m pytest tests/test_me_search_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
{
echo default; timeout 300 python tools/search_bench.py
for v in s3 s5 s6; do echo $v; B200RDO_LIB=$PWD/rav1e_b200/variants/libb200rdo_$v.so timeout 300 python tools/search_bench.py; done
} > gpurun_out/ab_search.log 2>&1
