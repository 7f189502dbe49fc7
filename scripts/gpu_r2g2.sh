# round 2, last session: full GPU suite after the capture fix (garbage collector collected before and disabled during
# every capture; thread-local capture mode).  Before: test_graphed_iteration_attack_matches_eager[mask-ce-bal-25-False]
# failed in the full run only -- the previous test's five CUDAGraph objects were collected in the middle of its capture.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x > gpurun_out/r2g2_pytest_gpu.log 2>&1); grep -n "FAILED\|passed\|failed" gpurun_out/r2g2_pytest_gpu.log | cut -c1-300
