for wl in "$@"; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python -c "
import json; r=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(r['value']), 'fps', round(r['ms_per_step'],3), 'ms; e2e', round(r['e2e']['value']), 'fps; top', r['roofline']['kernel'], r['roofline']['frac'], 'pipe', round(r['roofline']['pipeline_frac_of_peak'],4), 'Mpts/s', round(r['mpoints_per_sec']))" || tail -3 gpurun_out/bench_$wl.err
done
