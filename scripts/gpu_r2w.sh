# round 2, session 3, call 2: the reference's own PIR-AT Trainer under the drop-in (dropin.run_train_main)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --timeout 600 -x -k "run_train_main" > gpurun_out/r2w_train_main.log 2>&1); grep -n "^E \|^tests.*Error\|passed\|failed\|^train losses\|^evaluate (mAcc\|^parameters after" gpurun_out/r2w_train_main.log | cut -c1-600 | head -40
