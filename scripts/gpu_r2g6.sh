# round 2, last GPU seconds: the tests that use the oracle's up-sampling adjoint (rewritten as two matrix products)
(timeout 80 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_upsample.py -m gpu -q --timeout 60 -x -k "upsample" 2>&1 | tail -1)
