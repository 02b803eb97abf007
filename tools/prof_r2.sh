# final round-2 measurement set (1 GPU): full GPU test suite, smoke, bench line, per-kernel table, launch list, ncu capture
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
timeout 200 python bench.py --steps 5 --warmup 3 --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
timeout 900 python bench_kernels.py > gpurun_out/kernels_r2.jsonl 2> gpurun_out/kernels_r2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --workload 1080p --steps 2 --warmup 3 > gpurun_out/launches_r2_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:me_cand_group_u8 -s 3 -c 1 -f -o gpurun_out/prof_r2_final_sad python bench.py --workload 1080p --steps 2 --warmup 3 > gpurun_out/prof_r2_final.log 2>&1
