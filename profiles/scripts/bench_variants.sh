# usage: bench_variants.sh <lib> ...  — default-workload bench line (no CPU baseline) for the in-tree build and each variant
run() {
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bv_$1.json 2> gpurun_out/bv_$1.err
  python -c "
import json; r=json.load(open('gpurun_out/bv_$1.json')); k=r['roofline']['kernels_ms_per_step']; print('$1', round(r['ms_per_step'],3), round(r['e2e']['ms_per_step'],3), k)"
}
run main
cp online_3d_reconstruction_b200/libo3r.so /tmp/libo3r_keep.so
for v in "$@"; do cp online_3d_reconstruction_b200/$v online_3d_reconstruction_b200/libo3r.so; run $v; done
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
