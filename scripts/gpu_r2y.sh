# round 2, session 3, call 4: next-plane prefetch in the pow2 forward up-sampling kernels, blocks-per-SM sweep
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 200 -x -k "upsample or interpolate" 2>&1 | tail -3)
python scripts/up_fwd_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2y_up_fwd_probe.log
