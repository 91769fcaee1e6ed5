run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('$1', round(r['ms_per_step'],3), round(r['e2e']['ms_per_step'],3), 'vg', k.get('k_vg_reduce_w'), k.get('k_vg_reduce_s'), 'sort', k['k_rs_onesweep_u32'], 'emit', k['k_emit'])"; }
run default
O3R_GATHER_IN_SORT=1 run gather_in_sort
O3R_VG_SHORT=1 run vg_short
O3R_CARVEOUT_VG=100 run vg100
