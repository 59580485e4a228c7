set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_ordered.json 2> gpurun_out/bench_ordered.err; tail -c 3000 gpurun_out/bench_ordered.json; tail -5 gpurun_out/bench_ordered.err
python bench.py --steps 300 --warmup 20 --kernel vector --no-npb --no-cpu > gpurun_out/bench_vector.json 2> gpurun_out/bench_vector.err; tail -c 1500 gpurun_out/bench_vector.json
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu1.log 2>&1
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_stream_ordered -s 5 -c 2 -o gpurun_out/prof_ordered python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
