#!/bin/bash
# Final bench lines of a round: tools/bench_all.sh N "configs" tag   (run through gpurun [--gpus N])
N=${1:-1}; CFGS=${2:-"1 2 3 4 5"}; TAG=${3:-r02}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
PORT=29600
for c in $CFGS; do
  STEPS=20; [ $c -ge 4 ] && STEPS=10
  if [ "$N" = "1" ]; then
    python bench.py --config $c --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_n${N}_c$c.json 2> gpurun_out/${TAG}_bench_n${N}_c$c.err
  else
    PORT=$((PORT+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --config $c --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_n${N}_c$c.json 2> gpurun_out/${TAG}_bench_n${N}_c$c.err
  fi
  echo "config $c N=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_n${N}_c$c.err | tail -2
done
if [ "$N" = "1" ]; then
  python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_c2.json 2>/dev/null
  python bench.py --impl reference --config 3 --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_c3.json 2>/dev/null
  python bench.py --algo tc_full --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_bench_n1_c2_tc_full.json 2>/dev/null
fi
