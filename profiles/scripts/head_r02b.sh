# r02b evidence at HEAD (4 CTAs / SM tile kernel): default bench line, launch list, ncu --set full of k_tv
timeout 400 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02b_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_tv$ -s 2 -c 1 -o gpurun_out/r02b_tv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-steps 0 > gpurun_out/r02b_ncu2.log 2>&1
ls -la gpurun_out/r02b_*
