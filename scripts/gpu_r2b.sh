mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_graph_iteration.py "tests/test_gpu_parity.py::test_graphed_model_attack_equals_eager" -m gpu -q -x --timeout 300 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -80 > gpurun_out/r2b_pytest_graph.log); tail -70 gpurun_out/r2b_pytest_graph.log | cut -c1-400
(timeout 600 python -m pytest tests/test_gpu_e2e_rule.py tests/test_gpu_fused_upsample.py -m gpu -q -s --timeout 300 2>&1 | tail -30 > gpurun_out/r2b_pytest_rule.log); tail -25 gpurun_out/r2b_pytest_rule.log | cut -c1-400
