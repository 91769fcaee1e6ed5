# usage: sor_variants.sh <lib> ...   — SOR parity tests on the in-tree build, then the SOR bench line for it and each variant
timeout 300 python -m pytest tests -m gpu -x -q -k "sor or empty_and_tiny" 2>&1 | tail -3
run() {
  timeout 200 python bench.py --workload config2_semidense_720p_sor --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bs_$1.json 2> gpurun_out/bs_$1.err
  python -c "
import json; r=json.load(open('gpurun_out/bs_$1.json')); k=r['roofline']['kernels_ms_per_step']; print('$1', round(r['ms_per_step'],3), {a:b for a,b in k.items() if 'sor' in a})"
}
run main
cd online_3d_reconstruction_b200; cp libo3r.so /tmp/libo3r_keep.so; cd ..
for v in "$@"; do cp online_3d_reconstruction_b200/$v online_3d_reconstruction_b200/libo3r.so; run $v; done
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
