# round 2, last session: the up-sampling random sweep, then the full GPU suite on the final tree
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q --timeout 200 -k upsample > gpurun_out/r2g3_upfuzz.log 2>&1); grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2g3_upfuzz.log | cut -c1-260 | head -30
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x > gpurun_out/r2g3_pytest_gpu.log 2>&1); grep -n "^FAILED\|passed\|failed" gpurun_out/r2g3_pytest_gpu.log | cut -c1-300
