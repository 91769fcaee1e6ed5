# r02g: evidence of the final build: tile / fused / robustness / blur parity subset, default bench line, launch list, ncu --set full of
# k_tv and of the cooperative median kernel, the blur workloads
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_robustness.py tests/test_gpu_parity.py -x -q -k "fused or robust or blur or bilateral or tile or adjacent or prefetch" 2>&1 | tail -2 | tee gpurun_out/r02g_tests.log
timeout 400 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -c 300 gpurun_out/r02g_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02g_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_tv$ -s 2 -c 1 -o gpurun_out/r02g_tv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02g_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:^k_blur -s 1 -c 1 -o gpurun_out/r02g_median -f python bench.py --workload config2_semidense_720p_median31 --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02g_ncu5.log 2>&1
for w in config2_semidense_720p_blur30 config2_semidense_720p_median31 config4_sweep_720p_v002; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02g_bench_$w.json 2> gpurun_out/r02g_bench_$w.err; tail -c 200 gpurun_out/r02g_bench_$w.err
done
du -sh gpurun_out
