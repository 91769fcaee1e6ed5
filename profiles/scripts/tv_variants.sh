for v in ${VARIANTS:-libo3r_t512b512.so libo3r_t512b1024.so}; do
cd online_3d_reconstruction_b200; cp libo3r.so /tmp/libo3r_keep.so; cp $v libo3r.so; cd ..
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-steps 1 > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('$v', round(r['ms_per_step'],3), r['run']['partial_cells_per_step_per_gpu'], r['parity']['keys_equal'], r['parity']['counts_equal'], r['parity']['max_centroid_rel'], {a:k[a] for a in list(k)[:4]})"
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
done
