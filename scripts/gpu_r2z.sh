# round 2, session 3, call 5: strip / grid sweep of the pow2 backward up-sampling kernels
mkdir -p gpurun_out
python scripts/up_bwd_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2z_up_bwd_probe.log
