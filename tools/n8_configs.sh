# 8-GPU bench lines of the multi-GPU BASELINE configs (run through gpurun --gpus 8): tools/n8_configs.sh [c3 c4 c5]
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1"
port=29521
for c in "${@:-c3 c4 c5}"; do
  for cfg in $c; do
    steps=30; [ $cfg = c4 ] && steps=15; [ $cfg = c5 ] && steps=8
    $TR --master-port $port bench.py --gpus ${NGPU:-8} --config $cfg --steps $steps --warmup 4 --no-cpu > gpurun_out/n${NGPU:-8}_$cfg.json 2> gpurun_out/n${NGPU:-8}_$cfg.err
    port=$((port+1))
    echo "== $cfg"; grep '^{' gpurun_out/n${NGPU:-8}_$cfg.json | cut -c1-300; grep -i "error" gpurun_out/n${NGPU:-8}_$cfg.err | tail -2 | cut -c1-300
  done
done
