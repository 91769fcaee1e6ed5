# r02d: grouped 2-D copies of adjacent host planes: robustness tests + default bench line (e2e leg)
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_fused.py -x -q -k "adjacent or prefetch or cancelled or frame_mask" 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --parity-steps 1 > gpurun_out/r02d_bench_$i.json 2> gpurun_out/r02d_bench_$i.err; tail -c 300 gpurun_out/r02d_bench_$i.err
python -c "
import json; r=json.load(open('gpurun_out/r02d_bench_$i.json')); e=r['e2e']; print('run $i dev', round(r['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'compute-stream', round(e['compute_stream_ms_per_step'],3), 'floor', e['h2d_floor_ms_per_step'], e['h2d_bytes_per_step'], r['parity']['keys_equal'], r['parity']['records_equal'])"
done
for w in config4_sweep_720p_v002 config5_4k_u16_v001; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --parity-steps 0 > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err; tail -c 300 gpurun_out/r02d_bench_$w.err
python -c "
import json; r=json.load(open('gpurun_out/r02d_bench_$w.json')); e=r['e2e']; print('$w dev', round(r['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'floor', e['h2d_floor_ms_per_step'])"
done
