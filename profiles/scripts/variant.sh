# usage: variant.sh <variant .so name> [notest]; swaps the build in, runs a few parity tests + the bench, restores
cd online_3d_reconstruction_b200
cp libo3r.so /tmp/libo3r_keep.so
cp $1 libo3r.so
cd ..
if [ "$2" != "notest" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "cycles or voxel or tiled or frame_cloud" 2>&1 | tail -1; fi
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('$1', round(r['ms_per_step'],3), round(r['e2e']['ms_per_step'],3), k)"
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
