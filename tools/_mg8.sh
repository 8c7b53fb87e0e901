python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/g8_bench.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['parity_check'], d['e2e']['ms_per_step'])
for k,v in d['also'].items(): print(k, v['ms_per_step'], v['value'])
P
