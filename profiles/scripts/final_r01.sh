# final r01 evidence: full GPU parity suite, default + SOR bench lines, SOR launch list, ncu --set full of the SOR kernels
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_tests.log
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
timeout 300 python bench.py --workload config2_semidense_720p_sor --steps 3 --warmup 3 > gpurun_out/final_bench_sor.json 2> gpurun_out/final_bench_sor.err
python -c "
import json; r=json.load(open('gpurun_out/final_bench_sor.json')); print('sor', r['value'], r['ms_per_step'], r['e2e']['value'], r.get('cpu_baseline'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01m_sor.csv python bench.py --workload config2_semidense_720p_sor --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sor1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_sor_knn|k_sor_cells' -c 4 -o gpurun_out/prof_r01m_sor -f python bench.py --workload config2_semidense_720p_sor --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sor2.log 2>&1
ls -la gpurun_out/prof_r01m_sor.ncu-rep
