set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench57.json 2> gpurun_out/bench57.err; cut -c1-300 gpurun_out/bench57.json; tail -2 gpurun_out/bench57.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench57_ref.json 2>&1; cut -c1-200 gpurun_out/bench57_ref.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain57.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches57.csv python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu57.log 2>&1
timeout 300 python scripts/sweep.py D/8 "auto" 10 > gpurun_out/plain57c.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_panelr_kernel -s 5 -c 1 -o gpurun_out/prof_ring57 python scripts/sweep.py D/8 "auto" 10 > gpurun_out/ncu57c.log 2>&1
tail -2 gpurun_out/ncu57c.log
python -c "
import __graft_entry__ as e; e.smoke()"
