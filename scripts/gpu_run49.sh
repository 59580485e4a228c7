export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py C "panel,pr:G=2,pr:G=2;B=2,pr:G=4" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep49.txt
timeout 300 python scripts/sweep.py B "panel,pr:G=2,pr:G=2;B=2,pr:G=4" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep49.txt
timeout 300 python scripts/sweep.py A "panel,pr:G=2" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep49.txt
unset B200_SPMV_VERBOSE
timeout 600 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/bench49.json 2> gpurun_out/bench49.err; python -c "
import json; d=json.loads(open('gpurun_out/bench49.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"; tail -2 gpurun_out/bench49.err
