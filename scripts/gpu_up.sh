mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 150 -k "upsample or interpolate or hist or config1" 2>&1 | tail -12 > gpurun_out/pytest_up.log); tail -6 gpurun_out/pytest_up.log
python scripts/upsample_probe.py 2>&1 | grep -v Warn | tee gpurun_out/upsample_probe.log
python scripts/consumer_profile.py 16 150 all 2>&1 | head -3
