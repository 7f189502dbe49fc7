mkdir -p gpurun_out
python scripts/counts_probe.py 24 21 473 coherent 2>&1 | tee gpurun_out/r2m_counts_probe.log
python scripts/counts_probe.py 24 21 472 coherent 2>&1 | tee -a gpurun_out/r2m_counts_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_generic_ovf -c 14 -o gpurun_out/r2m_ovf_counts -f python scripts/counts_probe.py 24 21 473 coherent > gpurun_out/r2m_ncu1.log 2>&1; tail -2 gpurun_out/r2m_ncu1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_counts_launches.csv python scripts/counts_probe.py 24 21 473 coherent > /dev/null 2>&1; tail -12 gpurun_out/r2m_counts_launches.csv | cut -d, -f5,12- 
