# round 2, session 3, calls 5-7: strip / grid sweep of the pow2 backward up-sampling kernels; balanced plane groups
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 200 -x -k "upsample or interpolate" 2>&1 | tail -2)
python scripts/up_bwd_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2z_up_bwd_probe.log | grep default
python scripts/upsample_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2z_upsample_probe.log | cut -c1-110
