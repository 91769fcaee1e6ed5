# final r02 evidence (one GPU), build of the last commit: full GPU parity suite, default bench line + reference arm, launch list,
# ncu --set full of the tile kernel, the merge kernels and both blur kernels, N = 1 lines of the other configs and the blur workloads
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/r02e_tests.log
timeout 400 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02e_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_tv$ -s 2 -c 1 -o gpurun_out/r02e_tv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02e_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"k_acc|k_rs_onesweep|k_tv_compact|k_scan" -s 30 -c 24 -o gpurun_out/r02e_merge -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02e_ncu3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_bilateral$ -s 1 -c 1 -o gpurun_out/r02e_bilateral -f python bench.py --workload config2_semidense_720p_blur30 --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02e_ncu4.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_blur -s 1 -c 1 -o gpurun_out/r02e_median -f python bench.py --workload config2_semidense_720p_median31 --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02e_ncu5.log 2>&1
for w in config4_sweep_720p_v002 config5_4k_u16_v001 config3_dont_downsample_720p config1_sparse_720p config2_semidense_720p_blur30 config2_semidense_720p_median31; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02e_bench_$w.json 2> gpurun_out/r02e_bench_$w.err; tail -c 200 gpurun_out/r02e_bench_$w.err
done
python __graft_entry__.py smoke 2>&1 | tail -4
ls -la gpurun_out/r02e_*
