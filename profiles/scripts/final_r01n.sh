# final r01 state (atomicOr ranking): full GPU parity suite, default bench line, launch list, ncu --set full of the radix passes
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/final_tests_n.log
timeout 200 python bench.py > gpurun_out/final_bench_n.json 2> gpurun_out/final_bench_n.err; python -c "
import json; r=json.load(open('gpurun_out/final_bench_n.json')); print(r['value'], r['ms_per_step'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['kernels_ms_per_step'])"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_n1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_rs_onesweep -c 7 -o gpurun_out/prof_r01n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_n2.log 2>&1
ls -la gpurun_out/prof_r01n.ncu-rep gpurun_out/launches_r01n.csv
