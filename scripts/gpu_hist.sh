mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 -k "hist or metric or sea or iou" 2>&1 | tail -8 > gpurun_out/pytest_hist.log); tail -3 gpurun_out/pytest_hist.log
{
python scripts/hist_probe.py 64 150
python scripts/hist_probe.py 16 150
python scripts/hist_probe.py 48 21
echo "ROBSEG_CNT_PAIRS=4 ROBSEG_HIST_PAIRS=4"; ROBSEG_CNT_PAIRS=4 ROBSEG_HIST_PAIRS=4 python scripts/hist_probe.py 64 150
echo "ROBSEG_CNT_PER_SM=2"; ROBSEG_CNT_PER_SM=2 python scripts/hist_probe.py 64 150 > /tmp/o.txt; head -2 /tmp/o.txt
echo "ROBSEG_HIST_PER_SM=1"; ROBSEG_HIST_PER_SM=1 python scripts/hist_probe.py 64 150 > /tmp/o.txt; tail -3 /tmp/o.txt
} 2>&1 | grep -v Warning | tee gpurun_out/hist_probe.log
