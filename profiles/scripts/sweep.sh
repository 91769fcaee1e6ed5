run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('$1', round(r['ms_per_step'],3), round(r['e2e']['ms_per_step'],3), 'vg', k['k_vg_reduce_w'], 'acc', k['k_acc_reduce'], 'sort', k['k_rs_onesweep_u32'], 'emit', k['k_emit'])"; }
run default
O3R_CHUNK_FRAMES_DEV=25 run dev25
O3R_CHUNK_FRAMES_DEV=10 run dev10
O3R_CHUNK_FRAMES_DEV=5 run dev5
