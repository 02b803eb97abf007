# final round-1 measurement set (1 GPU): bench lines, per-kernel table, ncu capture + launch list
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 300 python bench.py --steps 10 --warmup 3 --pairs-per-launch 1 > gpurun_out/bench_r1_final_p1.json 2> gpurun_out/bench_r1_final_p1.err
timeout 900 python bench_kernels.py > gpurun_out/kernels_r1.jsonl 2> gpurun_out/kernels_r1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:me_cand_group_u8 -s 3 -c 1 -f -o gpurun_out/prof_r1_final python bench.py --steps 2 --warmup 3 > gpurun_out/prof_r1_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/launches_r1_final.log 2>&1
