set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench13.json 2> gpurun_out/bench13.err; cat gpurun_out/bench13.json; tail -3 gpurun_out/bench13.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench13_ref.json 2>&1; cat gpurun_out/bench13_ref.json
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain13.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu13.log 2>&1
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain13.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_panel_kernel -s 5 -c 1 -o gpurun_out/prof_panel13 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu13b.log 2>&1
tail -2 gpurun_out/ncu13b.log
