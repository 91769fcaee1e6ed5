# r02 scaling runs on the configs north_star names: bench.py under torch.distributed.run, N = $1 GPUs (one box)
# usage (on the GPU box): bash profiles/scripts/scale_r02.sh <N> <tag>
N=$1; TAG=$2
for w in config2_semidense_720p config4_sweep_720p_v002 config5_4k_u16_v001; do
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_${w}_${N}gpu.json 2> gpurun_out/${TAG}_${w}_${N}gpu.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w > gpurun_out/${TAG}_${w}_${N}gpu.json 2> gpurun_out/${TAG}_${w}_${N}gpu.err
  fi
  tail -c 200 gpurun_out/${TAG}_${w}_${N}gpu.err
  cut -c1-200 gpurun_out/${TAG}_${w}_${N}gpu.json
done
