set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench46.json 2> gpurun_out/bench46.err; cat gpurun_out/bench46.json | cut -c1-400; tail -2 gpurun_out/bench46.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench46_ref.json 2>&1; cut -c1-300 gpurun_out/bench46_ref.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain46.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches46.csv python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu46.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain46.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_panel_kernel -s 5 -c 1 -o gpurun_out/prof_panel46 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu46b.log 2>&1
tail -2 gpurun_out/ncu46b.log
timeout 300 python scripts/sweep.py D/8 "auto" 10 > gpurun_out/plain46c.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_panelr_kernel -s 5 -c 1 -o gpurun_out/prof_ring46 python scripts/sweep.py D/8 "auto" 10 > gpurun_out/ncu46c.log 2>&1
tail -2 gpurun_out/ncu46c.log
