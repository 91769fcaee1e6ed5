cd online_3d_reconstruction_b200
cp libo3r.so /tmp/libo3r_keep.so
for r in 32 64 96; do
  cp libo3r_r$r.so libo3r.so
  (cd ..; python bench.py --workload config2_semidense_720p_sor --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('rank $r', round(r['ms_per_step'],2), 'knn', k.get('k_sor_knn'), 'hard', k.get('k_sor_knn_hard'))")
done
cp /tmp/libo3r_keep.so libo3r.so
