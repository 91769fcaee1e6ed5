# r02i: grouped staging as 3-D copies (exact ROIs) instead of 2-D copies that carry the margin rows: grouped-copy tests + bench e2e
cd online_3d_reconstruction_b200; cp libo3r.so /tmp/libo3r_keep.so; cp libo3r_3d.so libo3r.so; cd ..
timeout 100 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_fused.py -x -q -k "adjacent or prefetch_and_chunked" 2>&1 | tail -2
timeout 100 python bench.py --no-cpu-baseline --parity-steps 1 > gpurun_out/r02i_bench_3d.json 2> gpurun_out/r02i_bench_3d.err; tail -c 200 gpurun_out/r02i_bench_3d.err
python -c "
import json; r=json.load(open('gpurun_out/r02i_bench_3d.json')); e=r['e2e']; print('3d dev', round(r['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), round(e['value']), 'floor', e['h2d_floor_ms_per_step'], r['parity']['keys_equal'], r['parity']['records_equal'])"
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
