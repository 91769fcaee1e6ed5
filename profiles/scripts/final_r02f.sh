# final r02 evidence, second half (the first call's 80 MB of ncu reports exceeded gpurun's 64 MiB return limit: its 145-test
# result is kept as profiles/r02_final_gpu_tests_call.log): default bench line + reference arm, launch list, ncu --set full of
# both blur kernels, N = 1 lines of the other configs and the blur workloads, A/B of the tabulated window offsets in k_tv
timeout 400 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 300 gpurun_out/r02f_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02f_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:^k_bilateral$ -s 1 -c 1 -o gpurun_out/r02f_bilateral -f python bench.py --workload config2_semidense_720p_blur30 --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02f_ncu4.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:^k_blur -s 1 -c 1 -o gpurun_out/r02f_median -f python bench.py --workload config2_semidense_720p_median31 --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02f_ncu5.log 2>&1
for w in config4_sweep_720p_v002 config5_4k_u16_v001 config3_dont_downsample_720p config1_sparse_720p config2_semidense_720p_blur30 config2_semidense_720p_median31; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02f_bench_$w.json 2> gpurun_out/r02f_bench_$w.err; tail -c 200 gpurun_out/r02f_bench_$w.err
done
cd online_3d_reconstruction_b200; cp libo3r.so /tmp/libo3r_keep.so; cd ..
for v in libo3r_offtab.so libo3r.so libo3r_offtab.so libo3r.so; do
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so; [ $v != libo3r.so ] && cp online_3d_reconstruction_b200/$v online_3d_reconstruction_b200/libo3r.so
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-steps 1 > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; r=json.load(open('gpurun_out/b.json')); k=r['roofline']['kernels_ms_per_step']; print('$v', round(r['ms_per_step'],4), r['parity']['keys_equal'], r['parity']['records_equal'], r['parity']['max_centroid_rel'], k['k_tv'])"
done
cp /tmp/libo3r_keep.so online_3d_reconstruction_b200/libo3r.so
du -sh gpurun_out
