# Single-GPU bench lines of the other BASELINE configs (run through gpurun): C1, C2, C4 with their CPU legs, C5 without
set -u
mkdir -p gpurun_out
for cfg in c1 c2 c4; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 > gpurun_out/n1_$cfg.json 2> gpurun_out/n1_$cfg.err
  echo "== $cfg"; grep '^{' gpurun_out/n1_$cfg.json | cut -c1-260; tail -1 gpurun_out/n1_$cfg.err | cut -c1-200
done
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/n1_c5.json 2> gpurun_out/n1_c5.err
echo "== c5"; grep '^{' gpurun_out/n1_c5.json | cut -c1-260; tail -1 gpurun_out/n1_c5.err | cut -c1-200
