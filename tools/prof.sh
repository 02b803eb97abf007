python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/tests_r1r.log
ncu --set full --clock-control none --import-source on -k regex:me_cand_group_u8 -s 3 -c 1 -f -o gpurun_out/prof_r1r python bench.py --steps 2 --warmup 3 > gpurun_out/prof_r1r.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1r.csv python bench.py --steps 2 --warmup 3 > gpurun_out/launches_r1r.log 2>&1
