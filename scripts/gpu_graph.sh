mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 -k "graphed or return_pred or early_stop" 2>&1 | tail -15 > gpurun_out/pytest_graph.log); tail -15 gpurun_out/pytest_graph.log
for g in "" "--graph"; do
  (timeout 600 python bench.py --batch 2 --classes 21 --no-cpu-baseline $g > gpurun_out/bench_b2$g.json 2> gpurun_out/bench_b2$g.err) ; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_b2$g.json')); print('B=2 C=21 $g', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/bench_b2$g.err
done
(timeout 600 python bench.py --no-cpu-baseline --graph > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err); python -c "
import json,sys; d=json.load(open('gpurun_out/bench_graph.json')); print('B=16 C=150 --graph', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/bench_graph.err
