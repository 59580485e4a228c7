set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cls in C A; do
B200_SPMV_ZEROCOPY=0 python scripts/e2e_probe.py $cls pinned 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=1 python scripts/e2e_probe.py $cls pinned 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=1 B200_SPMV_VALIDATE=0 python scripts/e2e_probe.py $cls pinned 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=1 B200_SPMV_TIME_KERNELS=0 python scripts/e2e_probe.py $cls pinned 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=0 python scripts/e2e_probe.py $cls pageable 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=1 python scripts/e2e_probe.py $cls pageable 2>&1 | tee -a gpurun_out/e2e6.txt
B200_SPMV_ZEROCOPY=1 B200_SPMV_PIN_HOST=1 python scripts/e2e_probe.py $cls pageable 2>&1 | tee -a gpurun_out/e2e6.txt
done
python bench.py --steps 500 --warmup 20 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cat gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err
