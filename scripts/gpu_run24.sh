set -x
python -m pytest tests -m gpu -x -q -k "kat or golden_reference or ragged or around_the_tile or merge_kernel or sell_long" 2>&1 | tail -3 &&
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck24.log python -m pytest tests -m gpu -x -q -k "kat or golden_reference or ragged or around_the_tile or merge_kernel or sell_long" > gpurun_out/memcheck24_pytest.log 2>&1
echo sanitizer_rc=$?
tail -3 gpurun_out/memcheck24_pytest.log; tail -5 gpurun_out/memcheck24.log
