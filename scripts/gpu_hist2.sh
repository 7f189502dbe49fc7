mkdir -p gpurun_out
{
for n in 3 6 9 12 24; do echo "ROBSEG_CNT_PER_SM=$n"; ROBSEG_CNT_PER_SM=$n python scripts/hist_probe.py 64 150 > /tmp/o.txt; head -2 /tmp/o.txt;  ROBSEG_CNT_PER_SM=$n python scripts/hist_probe.py 16 150 > /tmp/o.txt; head -2 /tmp/o.txt; done
for n in 2 4 6 12; do echo "ROBSEG_HIST_PER_SM=$n"; ROBSEG_HIST_PER_SM=$n python scripts/hist_probe.py 64 150 > /tmp/o.txt; tail -3 /tmp/o.txt; ROBSEG_HIST_PER_SM=$n python scripts/hist_probe.py 16 150 > /tmp/o.txt; tail -3 /tmp/o.txt; done
python scripts/hist_probe.py 256 150 | head -4
} 2>&1 | grep -v Warning | tee gpurun_out/hist_probe2.log
