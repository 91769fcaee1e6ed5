# r02c: the rewritten median (shared-memory atomics, tracked median) and bilateral (two pixels per thread) kernels:
# parity tests, then the two blur workloads of bench.py
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "blur or bilateral" 2>&1 | tail -3
for w in config2_semidense_720p_blur30 config2_semidense_720p_median31; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --parity-steps 1 > gpurun_out/r02c_bench_$w.json 2> gpurun_out/r02c_bench_$w.err; tail -c 300 gpurun_out/r02c_bench_$w.err
python -c "
import json; r=json.load(open('gpurun_out/r02c_bench_$w.json')); print('$w', round(r['ms_per_step'],3), round(r['e2e']['ms_per_step'],3), r['roofline']['kernels_ms_per_step'], r['parity'])"
done
