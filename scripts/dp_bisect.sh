for cfg in "side_wgrad=0,fork_frozen=0" "side_wgrad=0" "fork_frozen=0"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --cfg $cfg > gpurun_out/bis.json 2> gpurun_out/bis.err; echo "cfg=$cfg rc=$? $(tail -c 200 gpurun_out/bis.json | head -c 0) $(grep -c 'dependency created' gpurun_out/bis.err) $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/bis.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['dp_check'])
except Exception as e: print('noline')")"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --dp-overlap 0 > gpurun_out/bis.json 2> gpurun_out/bis.err; echo "overlap0 rc=$? $(grep -c 'dependency created' gpurun_out/bis.err)"; tail -c 300 gpurun_out/bis.json
